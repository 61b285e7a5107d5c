// sffg_double.cpp -- TEST DOUBLE of the engine's C ABI (include/sffg.h).  TEST INFRASTRUCTURE ONLY.
//
// The batched planner hosts (space_filling_forest_star_b200/host/) sit strictly above the C ABI, so their host-side
// logic -- candidate generation, replay of the reference's accept / rewire / merge rules, path extraction -- can be
// exercised without a GPU by linking them against this double instead of libsffg.so.  Every compute entry point here is
// answered by the CPU oracle (oracle/sff_oracle.c: RAPID-style OBB-tree verdicts, reference edge sampling, exact linear
// k-NN / radius).  It is built only by tests/test_planner_host_cpu.py into oracle/_build/ and is never loaded,
// linked or shipped by the product: libsffg.so itself has no CPU path and refuses to run without an sm_100 device.
// Only the subset of the ABI the planner hosts call is provided.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../space_filling_forest_star_b200/csrc/common.h"

extern "C" {
// oracle/sff_oracle.c
struct orc_model;
orc_model *orc_model_build(const double *tris, int n);
void orc_model_free(orc_model *m);
void orc_collide_obbtree_batch(const orc_model *obst, const orc_model *robot, const double *poses, int64_t n, int first_contact,
                               uint8_t *verdicts, int64_t *counters, int threads);
void orc_edge_free_batch(const double *obst, int nT, const double *robot, int nR, const orc_model *mo, const orc_model *mr,
                         const double *starts, const double *ends, int64_t m, double sample, int rot_mode, uint8_t *free_out,
                         int32_t *first_hit_out, int64_t *samples_tested, int threads);
void orc_knn_linear(const float *nodes, int64_t n, int dim, const float *queries, int64_t nq, int k, int32_t *ids_out, float *d2_out,
                    int threads);
void orc_radius_linear(const float *nodes, int64_t n, int dim, const float *queries, int64_t nq, float r2, int32_t *counts,
                       const int64_t *offsets, int32_t *ids_out, float *d2_out, int threads);
}

namespace sffg {
static thread_local std::string g_err;
void set_error(const std::string &msg) { g_err = msg; }
int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}
}  // namespace sffg

struct sffg_env {
  std::vector<double> obst, robot;
  orc_model *mo = nullptr, *mr = nullptr;
};
struct sffg_index {
  int dim;
  std::vector<float> rows;
};

extern "C" {

SFFG_API int sffg_version(void) { return -1; }   // negative: not the engine
SFFG_API const char *sffg_last_error(void) { return sffg::g_err.c_str(); }
SFFG_API int sffg_init(int) { return SFFG_OK; }

SFFG_API int sffg_mesh_load(const char *path, int is_obj, const double position[3], double scale, double **tris_out,
                            int64_t *n_tris_out, double bbox_out[6]) {
  const double zero[3] = {0, 0, 0};
  std::vector<double> tris;
  double bbox[6];
  int rc = sffg::load_mesh(path, is_obj, position ? position : zero, scale, &tris, bbox);
  if (rc != SFFG_OK) return rc;
  double *out = (double *)std::malloc(std::max<size_t>(tris.size(), 1) * sizeof(double));
  std::memcpy(out, tris.data(), tris.size() * sizeof(double));
  *tris_out = out;
  *n_tris_out = (int64_t)(tris.size() / 9);
  if (bbox_out) std::memcpy(bbox_out, bbox, sizeof bbox);
  return SFFG_OK;
}
SFFG_API void sffg_free(void *p) { std::free(p); }

SFFG_API int sffg_env_create(const double *obst_tris, int64_t n_obst, const double *robot_tris, int64_t n_robot, sffg_env **out) {
  sffg_env *e = new sffg_env;
  if (n_obst) e->obst.assign(obst_tris, obst_tris + 9 * n_obst);
  e->robot.assign(robot_tris, robot_tris + 9 * n_robot);
  if (n_obst) e->mo = orc_model_build(e->obst.data(), (int)n_obst);
  e->mr = orc_model_build(e->robot.data(), (int)n_robot);
  *out = e;
  return SFFG_OK;
}
SFFG_API int sffg_env_destroy(sffg_env *e) {
  if (e) {
    orc_model_free(e->mo);
    orc_model_free(e->mr);
    delete e;
  }
  return SFFG_OK;
}

SFFG_API int sffg_collide_poses_f64(sffg_env *e, const double *poses, int64_t n, uint8_t *verdict_out) {
  if (!e->mo) {
    std::memset(verdict_out, 0, (size_t)n);
    return SFFG_OK;
  }
  orc_collide_obbtree_batch(e->mo, e->mr, poses, n, 1, verdict_out, nullptr, 0);
  return SFFG_OK;
}

SFFG_API int sffg_check_edges(sffg_env *e, const double *starts, const double *ends, int64_t m, double sample_dist, int rot_mode,
                              uint8_t *free_out, int32_t *first_hit_out) {
  if (!e->mo) {
    std::memset(free_out, 1, (size_t)m);
    if (first_hit_out) std::memset(first_hit_out, 0, (size_t)m * sizeof(int32_t));
    return SFFG_OK;
  }
  orc_edge_free_batch(e->obst.data(), (int)(e->obst.size() / 9), e->robot.data(), (int)(e->robot.size() / 9), e->mo, e->mr, starts,
                      ends, m, sample_dist, rot_mode, free_out, first_hit_out, nullptr, 0);
  return SFFG_OK;
}

SFFG_API int sffg_check_moves(sffg_env *e, const double *starts, const double *ends, int64_t m, double sample_dist, int rot_mode,
                              uint8_t *ok_out) {
  std::vector<uint8_t> hit((size_t)m);
  sffg_collide_poses_f64(e, ends, m, hit.data());
  sffg_check_edges(e, starts, ends, m, sample_dist, rot_mode, ok_out, nullptr);
  for (int64_t i = 0; i < m; ++i) ok_out[i] = (uint8_t)(ok_out[i] && !hit[(size_t)i]);
  return SFFG_OK;
}

SFFG_API int sffg_index_create(int dim, sffg_index **out) {
  if (dim != 2 && dim != 6) return sffg::fail(SFFG_ERR_ARG, "dim must be 2 or 6");
  *out = new sffg_index{dim, {}};
  return SFFG_OK;
}
SFFG_API int sffg_index_destroy(sffg_index *idx) {
  delete idx;
  return SFFG_OK;
}
SFFG_API int sffg_index_add(sffg_index *idx, const float *pts, int64_t n) {
  idx->rows.insert(idx->rows.end(), pts, pts + n * idx->dim);
  return SFFG_OK;
}
SFFG_API int sffg_index_add_multi(sffg_index *const *idx, const int64_t *n_per, int n_idx, const float *pts) {
  int64_t off = 0;
  for (int i = 0; i < n_idx; ++i) {
    sffg_index_add(idx[i], pts + off * idx[i]->dim, n_per[i]);
    off += n_per[i];
  }
  return SFFG_OK;
}
SFFG_API int64_t sffg_index_size(const sffg_index *idx) { return (int64_t)(idx->rows.size() / idx->dim); }

SFFG_API int sffg_knn(sffg_index *idx, const float *queries, int64_t nq, int k, int32_t *ids_out, float *d2_out) {
  if (k < 1 || k > SFFG_MAX_K) return sffg::fail(SFFG_ERR_ARG, "k out of range");
  orc_knn_linear(idx->rows.data(), sffg_index_size(idx), idx->dim, queries, nq, k, ids_out, d2_out, 0);
  return SFFG_OK;
}
SFFG_API int sffg_knn_multi(sffg_index *const *idx, const int64_t *nq_per, int n_idx, const float *queries, int k, int32_t *ids_out,
                            float *d2_out) {
  int64_t row = 0;
  for (int i = 0; i < n_idx; ++i) {
    if (nq_per[i] == 0) continue;
    int rc = sffg_knn(idx[i], queries + row * idx[i]->dim, nq_per[i], k, ids_out + row * k, d2_out + row * k);
    if (rc != SFFG_OK) return rc;
    row += nq_per[i];
  }
  return SFFG_OK;
}

SFFG_API int sffg_radius(sffg_index *idx, const float *queries, int64_t nq, float r2, int32_t *counts_out, int32_t *ids_out,
                         float *d2_out, int64_t capacity, int64_t *total_out) {
  const int64_t n = sffg_index_size(idx);
  orc_radius_linear(idx->rows.data(), n, idx->dim, queries, nq, r2, counts_out, nullptr, nullptr, nullptr, 0);
  std::vector<int64_t> off((size_t)nq + 1, 0);
  for (int64_t i = 0; i < nq; ++i) off[i + 1] = off[i] + counts_out[i];
  if (total_out) *total_out = off[nq];
  if (!ids_out) return SFFG_OK;
  if (capacity < off[nq]) return sffg::fail(SFFG_ERR_CAPACITY, "radius result buffer too small");
  orc_radius_linear(idx->rows.data(), n, idx->dim, queries, nq, r2, counts_out, off.data(), ids_out, d2_out, 0);
  return SFFG_OK;
}

// asynchronous forms: the double answers at once, the end calls have nothing left to do
SFFG_API int sffg_radius_begin(sffg_index *idx, const float *queries, int64_t nq, float r2, int32_t *counts_out, int32_t *ids_out,
                               float *d2_out, int64_t capacity, int64_t *total_out) {
  return sffg_radius(idx, queries, nq, r2, counts_out, ids_out, d2_out, capacity, total_out);
}
SFFG_API int sffg_knn_multi_begin(sffg_index *const *idx, const int64_t *nq_per, int n_idx, const float *queries, int k,
                                  int32_t *ids_out, float *d2_out) {
  return sffg_knn_multi(idx, nq_per, n_idx, queries, k, ids_out, d2_out);
}
SFFG_API int sffg_index_add_multi_begin(sffg_index *const *idx, const int64_t *n_per, int n_idx, const float *pts) {
  return sffg_index_add_multi(idx, n_per, n_idx, pts);
}
SFFG_API int sffg_index_end(sffg_index *) { return SFFG_OK; }
SFFG_API int sffg_check_edges_begin(sffg_env *e, const double *starts, const double *ends, int64_t m, double sample_dist, int rot_mode,
                                    uint8_t *free_out, int32_t *first_hit_out) {
  return sffg_check_edges(e, starts, ends, m, sample_dist, rot_mode, free_out, first_hit_out);
}
SFFG_API int sffg_check_moves_begin(sffg_env *e, const double *starts, const double *ends, int64_t m, double sample_dist, int rot_mode,
                                    uint8_t *ok_out) {
  return sffg_check_moves(e, starts, ends, m, sample_dist, rot_mode, ok_out);
}
SFFG_API int sffg_env_end(sffg_env *) { return SFFG_OK; }

}  // extern "C"
