// ref_flann.cpp -- thin C entry points around the REAL vendored FLANN 1.9.1 of the reference.
//
// TEST INFRASTRUCTURE ONLY (see oracle/sff_oracle.c header).  Built by `make -C oracle ref` into
// oracle/_ref/libflann_ref.so from the reference's own sources where they lie under /root/reference
// (nothing is copied): <flann/flann.hpp> (lib/flann/src/cpp) and src/primitives.h (D6Distance, AngleDifference).
//
// Two uses:
//   * ref_linear_*  : flann::LinearIndex (algorithms/linear_index.h:132-147, pinned exact by FLANN's own
//                     flann_linear_test.cpp:20-24) with the metric bug of src/primitives.h:418/:423 fixed
//                     (`+=`, size-bounded loops).  This pins oracle/sff_oracle.c::orc_knn_linear/orc_radius_linear.
//   * ref_planner_* : the index exactly as the planner builds and queries it (KDTreeIndexParams(4), one point
//                     at a time via addPoints, SearchParams(128), ORIGINAL D6Distance) -- src/forest.h:72-73,
//                     :266, :317, :367.  Speed baseline only ("reference RAPID+FLANN CPU path").
#include <cstdint>
#include <cstring>
#include <vector>
#include "primitives.h"   // from /root/reference/src : D6Distance<T>, AngleDifference, NormalizeAngle

namespace {

// D6Distance as intended: accumulate, and honour `size` (2-D rows hold only 2 floats).
template <class T>
struct FixedD6 {
  typedef bool is_kdtree_distance;
  typedef T ElementType;
  typedef typename flann::Accumulator<T>::Type ResultType;

  template <typename It1, typename It2>
  ResultType operator()(It1 a, It2 b, size_t size, ResultType = -1) const {
    ResultType result = ResultType();
    ResultType diff;
    size_t lin = size < 3 ? size : 3;
    for (size_t i = 0; i < lin; ++i) {
      diff = (ResultType)(*a++ - *b++);
      result += diff * diff;
    }
    for (size_t i = 3; i < size; ++i) {
      diff = (ResultType)AngleDifference(*a++, *b++);
      result += diff * diff;
    }
    return result;
  }
  template <typename U, typename V>
  inline ResultType accum_dist(const U &a, const V &b, int part) const {
    if (part > 2) {
      ResultType d = (ResultType)AngleDifference(a, b);
      return d * d;
    }
    return (a - b) * (a - b);
  }
};

template <class Index>
void copy_rows(const std::vector<std::vector<int>> &ind, const std::vector<std::vector<float>> &dst, int64_t nq, int k,
               int32_t *ids_out, float *d2_out) {
  for (int64_t q = 0; q < nq; ++q)
    for (int j = 0; j < k; ++j) {
      bool ok = j < (int)ind[q].size();
      ids_out[q * k + j] = ok ? ind[q][j] : -1;
      d2_out[q * k + j] = ok ? dst[q][j] : INFINITY;
    }
}

struct PlannerIndex {
  flann::Index<D6Distance<float>> *idx = nullptr;
  std::vector<float *> rows;   // the planner keeps every added row alive (Tree::ptrToDel, src/primitives.h:507)
  int dim = 0;
  ~PlannerIndex() {
    delete idx;
    for (float *p : rows) delete[] p;
  }
};

}  // namespace

extern "C" {

__attribute__((visibility("default"))) int ref_linear_knn(const float *nodes, int64_t n, int dim, const float *queries,
                                                         int64_t nq, int k, int32_t *ids_out, float *d2_out, int cores) {
  flann::Matrix<float> data(const_cast<float *>(nodes), n, dim);
  flann::Matrix<float> q(const_cast<float *>(queries), nq, dim);
  flann::Index<FixedD6<float>> index(data, flann::LinearIndexParams());
  index.buildIndex();
  std::vector<std::vector<int>> ind;
  std::vector<std::vector<float>> dst;
  flann::SearchParams sp(flann::FLANN_CHECKS_UNLIMITED);
  sp.cores = cores;
  index.knnSearch(q, ind, dst, (size_t)k, sp);
  copy_rows<void>(ind, dst, nq, k, ids_out, d2_out);
  return 0;
}

// counts[nq] always written.  If ids_out != NULL, rows go to offsets[q] (caller's exclusive scan of counts).
__attribute__((visibility("default"))) int ref_linear_radius(const float *nodes, int64_t n, int dim, const float *queries,
                                                            int64_t nq, float r2, int32_t *counts, const int64_t *offsets,
                                                            int32_t *ids_out, float *d2_out, int cores) {
  flann::Matrix<float> data(const_cast<float *>(nodes), n, dim);
  flann::Matrix<float> q(const_cast<float *>(queries), nq, dim);
  flann::Index<FixedD6<float>> index(data, flann::LinearIndexParams());
  index.buildIndex();
  std::vector<std::vector<int>> ind;
  std::vector<std::vector<float>> dst;
  flann::SearchParams sp(flann::FLANN_CHECKS_UNLIMITED);
  sp.cores = cores;
  index.radiusSearch(q, ind, dst, r2, sp);
  for (int64_t i = 0; i < nq; ++i) {
    counts[i] = (int32_t)ind[i].size();
    if (ids_out)
      for (size_t j = 0; j < ind[i].size(); ++j) {
        ids_out[offsets[i] + j] = ind[i][j];
        d2_out[offsets[i] + j] = dst[i][j];
      }
  }
  return 0;
}

// ---- the planner's own index configuration (approximate + original metric) -------------------------------
__attribute__((visibility("default"))) void *ref_planner_index_build(const float *nodes, int64_t n, int dim) {
  PlannerIndex *p = new PlannerIndex();
  p->dim = dim;
  for (int64_t i = 0; i < n; ++i) {
    float *row = new float[dim];
    std::memcpy(row, nodes + i * dim, sizeof(float) * dim);
    p->rows.push_back(row);
    flann::Matrix<float> m(row, 1, dim);
    if (i == 0) {
      p->idx = new flann::Index<D6Distance<float>>(m, flann::KDTreeIndexParams(4));   // src/forest.h:72
      p->idx->buildIndex();                                                           // src/forest.h:73
    } else {
      p->idx->addPoints(m);                                                           // src/forest.h:367
    }
  }
  return p;
}
__attribute__((visibility("default"))) void ref_planner_index_free(void *h) { delete (PlannerIndex *)h; }

__attribute__((visibility("default"))) int ref_planner_knn(void *h, const float *queries, int64_t nq, int k,
                                                          int32_t *ids_out, float *d2_out, int cores) {
  PlannerIndex *p = (PlannerIndex *)h;
  std::vector<std::vector<int>> ind;
  std::vector<std::vector<float>> dst;
  flann::SearchParams sp(128);   // src/forest.h:317
  sp.cores = cores;
  if (cores == 1) {              // one query per call, as the planner does
    for (int64_t q = 0; q < nq; ++q) {
      flann::Matrix<float> m(const_cast<float *>(queries + q * p->dim), 1, p->dim);
      ind.clear(); dst.clear();
      p->idx->knnSearch(m, ind, dst, (size_t)k, sp);
      for (int j = 0; j < k; ++j) {
        bool ok = j < (int)ind[0].size();
        ids_out[q * k + j] = ok ? ind[0][j] : -1;
        d2_out[q * k + j] = ok ? dst[0][j] : INFINITY;
      }
    }
  } else {
    flann::Matrix<float> m(const_cast<float *>(queries), nq, p->dim);
    p->idx->knnSearch(m, ind, dst, (size_t)k, sp);
    copy_rows<void>(ind, dst, nq, k, ids_out, d2_out);
  }
  return 0;
}

__attribute__((visibility("default"))) int64_t ref_planner_radius(void *h, const float *queries, int64_t nq, float r2,
                                                                 int32_t *counts, int cores) {
  PlannerIndex *p = (PlannerIndex *)h;
  std::vector<std::vector<int>> ind;
  std::vector<std::vector<float>> dst;
  flann::SearchParams sp(128);   // src/forest.h:266-267
  sp.cores = cores;
  int64_t total = 0;
  if (cores == 1) {
    for (int64_t q = 0; q < nq; ++q) {
      flann::Matrix<float> m(const_cast<float *>(queries + q * p->dim), 1, p->dim);
      ind.clear(); dst.clear();
      int c = p->idx->radiusSearch(m, ind, dst, r2, sp);
      if (counts) counts[q] = c;
      total += c;
    }
  } else {
    flann::Matrix<float> m(const_cast<float *>(queries), nq, p->dim);
    total = p->idx->radiusSearch(m, ind, dst, r2, sp);
    if (counts) for (int64_t q = 0; q < nq; ++q) counts[q] = (int32_t)ind[q].size();
  }
  return total;
}

}  // extern "C"
