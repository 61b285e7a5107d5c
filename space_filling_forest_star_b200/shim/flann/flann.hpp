// flann/flann.hpp -- drop-in shim: the subset of FLANN 1.9.1's C++ API the planner uses (reference src/forest.h:59-100,
// :257-267, :313-317, :367; src/rrt.h:53-77, :139-166, :209-298; src/lazy.h:168-259; src/primitives.h:404-438, :506),
// forwarded to the B200 engine's exact neighbour index through the C ABI of include/sffg.h.
//
// Put this directory in front of lib/flann/src/cpp on the include path and link libsffg.so: forest.h / rrt.h / lazy.h /
// primitives.h compile unchanged.  Semantics kept from flann::Index (lib/flann/src/cpp/flann/flann.hpp:75-368):
// ids are insertion order, distances are squared, knnSearch rows ascend, radiusSearch is strict (d2 < r2) and sorted.
// Deliberate difference: the search is EXACT with the intended metric (the Distance functor is only consulted for its
// element types), not the 4-kd-tree / 128-checks approximation with the assigning D6Distance.
#ifndef SFFG_FLANN_SHIM_HPP_
#define SFFG_FLANN_SHIM_HPP_

#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "sffg.h"

namespace flann {

class FLANNException : public std::runtime_error {
 public:
  explicit FLANNException(const std::string &m) : std::runtime_error(m) {}
};

template <typename T> struct Accumulator { typedef T Type; };
template <> struct Accumulator<unsigned char> { typedef float Type; };
template <> struct Accumulator<char> { typedef float Type; };
template <> struct Accumulator<short> { typedef float Type; };
template <> struct Accumulator<int> { typedef float Type; };

// non-owning row-major matrix (util/matrix.h:111-132)
template <typename T>
class Matrix {
 public:
  size_t rows = 0, cols = 0, stride = 0;
  Matrix() {}
  Matrix(T *data, size_t r, size_t c, size_t stride_ = 0) : rows(r), cols(c), stride(stride_ ? stride_ : c * sizeof(T)), data_(data) {}
  T *operator[](size_t i) const { return reinterpret_cast<T *>(reinterpret_cast<unsigned char *>(data_) + i * stride); }
  T *ptr() const { return data_; }

 private:
  T *data_ = nullptr;
};

struct IndexParams {};
struct KDTreeIndexParams : IndexParams { explicit KDTreeIndexParams(int = 4) {} };
struct LinearIndexParams : IndexParams {};
const int FLANN_CHECKS_UNLIMITED = -1;
struct SearchParams {
  explicit SearchParams(int checks_ = 32, float eps_ = 0, bool sorted_ = true) : checks(checks_), eps(eps_), sorted(sorted_) {}
  int checks;
  float eps;
  bool sorted;
  int max_neighbors = -1;
  int cores = 1;
};

template <typename Distance>
class Index {
 public:
  typedef typename Distance::ElementType ElementType;
  typedef typename Distance::ResultType DistanceType;

  Index(const Matrix<ElementType> &points, const IndexParams &, Distance = Distance()) : initial_(points) {
    dim_ = (int)points.cols;
    int rc = sffg_index_create(dim_, &idx_);
    if (rc != SFFG_OK) throw FLANNException(sffg_last_error());
  }
  ~Index() { sffg_index_destroy(idx_); }
  Index(const Index &) = delete;
  Index &operator=(const Index &) = delete;

  void buildIndex() {
    if (built_) return;
    add(initial_);
    built_ = true;
  }
  void addPoints(const Matrix<ElementType> &points, float /*rebuild_threshold*/ = 2) { add(points); }
  size_t size() const { return (size_t)sffg_index_size(idx_); }
  size_t veclen() const { return (size_t)dim_; }

  int knnSearch(const Matrix<ElementType> &queries, std::vector<std::vector<int>> &indices,
                std::vector<std::vector<DistanceType>> &dists, size_t knn, const SearchParams &) const {
    const size_t nq = queries.rows;
    indices.assign(nq, std::vector<int>());
    dists.assign(nq, std::vector<DistanceType>());
    if (knn == 0 || nq == 0) return 0;
    if (knn > SFFG_MAX_K) knn = SFFG_MAX_K;
    std::vector<float> q = pack(queries);
    std::vector<int32_t> ids(nq * knn);
    std::vector<float> d2(nq * knn);
    if (sffg_knn(idx_, q.data(), (int64_t)nq, (int)knn, ids.data(), d2.data()) != SFFG_OK) throw FLANNException(sffg_last_error());
    int count = 0;
    for (size_t i = 0; i < nq; ++i)
      for (size_t j = 0; j < knn && ids[i * knn + j] >= 0; ++j) {
        indices[i].push_back(ids[i * knn + j]);
        dists[i].push_back((DistanceType)d2[i * knn + j]);
        ++count;
      }
    return count;
  }

  int radiusSearch(const Matrix<ElementType> &queries, std::vector<std::vector<int>> &indices,
                   std::vector<std::vector<DistanceType>> &dists, float radius, const SearchParams &) const {
    const size_t nq = queries.rows;
    indices.assign(nq, std::vector<int>());
    dists.assign(nq, std::vector<DistanceType>());
    if (nq == 0) return 0;
    std::vector<float> q = pack(queries);
    std::vector<int32_t> counts(nq);
    int64_t total = 0;
    if (sffg_radius(idx_, q.data(), (int64_t)nq, radius, counts.data(), nullptr, nullptr, 0, &total) != SFFG_OK)
      throw FLANNException(sffg_last_error());
    if (total == 0) return 0;
    std::vector<int32_t> ids((size_t)total);
    std::vector<float> d2((size_t)total);
    if (sffg_radius(idx_, q.data(), (int64_t)nq, radius, counts.data(), ids.data(), d2.data(), total, &total) != SFFG_OK)
      throw FLANNException(sffg_last_error());
    size_t off = 0;
    for (size_t i = 0; i < nq; ++i) {
      indices[i].assign(ids.begin() + off, ids.begin() + off + counts[i]);
      dists[i].assign(d2.begin() + off, d2.begin() + off + counts[i]);
      off += (size_t)counts[i];
    }
    return (int)total;
  }

 private:
  std::vector<float> pack(const Matrix<ElementType> &m) const {
    std::vector<float> out(m.rows * (size_t)dim_);
    for (size_t i = 0; i < m.rows; ++i)
      for (int c = 0; c < dim_; ++c) out[i * dim_ + c] = (float)m[i][c];
    return out;
  }
  void add(const Matrix<ElementType> &m) {
    if (m.rows == 0) return;
    if ((int)m.cols != dim_) throw FLANNException("addPoints: dimension mismatch");
    std::vector<float> p = pack(m);
    if (sffg_index_add(idx_, p.data(), (int64_t)m.rows) != SFFG_OK) throw FLANNException(sffg_last_error());
  }
  Matrix<ElementType> initial_;
  sffg_index *idx_ = nullptr;
  int dim_ = 0;
  bool built_ = false;
};

}  // namespace flann

#endif  // SFFG_FLANN_SHIM_HPP_
