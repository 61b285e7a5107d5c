// collide_kernels.cu -- sm_100a kernels for pose and edge collision verdicts.
//
// Replaces the inside of Environment<T>::Collide -> Obstacle<T>::Collide -> RAPID_Collide
// (reference src/environment.h:306-316, :269-276) and of Solver<T,R>::isPathFree (src/problemStruct.h:154-168).
//
// Execution model (one warp = one pose at a time):
//   phase A  lane-per-pose: load 32 poses, cull the robot's bounding sphere against the obstacle AABB, build R
//   phase B  warp-per-pose for the survivors:
//            - 8-wide AABB BVH (breadth-first node order); the top of the hierarchy -- a <=32-box cut tested by all lanes
//              in step 0 and the first three levels of nodes -- is staged in shared memory once per CTA; warp-shared DFS
//              stack in shared memory; each step pops up to 4 nodes and the 32 lanes test the 4x8 child boxes against
//              the robot's oriented box (6-axis conservative SAT)
//            - surviving leaf triangles are transformed into the robot frame (as RAPID does) lane-per-triangle,
//              then lane-per-(obstacle triangle, robot triangle) pair runs the 17-axis separating-axis test in FP32
//              as a *certificate* test: an axis only counts when its gap exceeds a rigorous rounding bound
//            - pairs the lane-per-pair stage leaves open get the 9 edge x edge axes and, only if those fail, the 6
//              contact certificates cooperatively, three pairs per pass on ten lanes each
//            - pairs without an FP32 certificate are decided by the exact stage: the same 17 axes in FP64 with the
//              operation order of the CPU oracle, one axis per lane; __all_sync gives the pair verdict
//            - __ballot/__any early-out on the first confirmed contact (verdict-equivalent to RAPID's ALL_CONTACTS
//              because the caller only looks at num_contacts != 0)
// The result equals "OR over all triangle pairs of the double-precision SAT" -- the oracle's ground truth.
#include <cstdint>
#include <mutex>

#include "collide_kernels.cuh"

namespace sffg {
namespace {

#ifndef SFFG_WARPS
#define SFFG_WARPS 16
#endif
#ifndef SFFG_TRI_FLUSH
#define SFFG_TRI_FLUSH 1
#endif
constexpr int kWarpsPerBlock = SFFG_WARPS;
constexpr int kThreads = kWarpsPerBlock * 32;
constexpr int kStackCap = 384;   // node ids pending for one pose
constexpr int kTriCap = 64;      // candidate triangles pending for one pose
constexpr int kTriFlush = SFFG_TRI_FLUSH;    // run the triangle stage once this many candidates are pending
constexpr unsigned kFull = 0xffffffffu;
#ifndef SFFG_MIN_BLOCKS
#define SFFG_MIN_BLOCKS 1
#endif
constexpr float kEpsBox = 1.52587890625e-05f;   // 2^-16, relative slack of the box culls   (>= 140 ulp)
constexpr float kEpsSat = 1.52587890625e-05f;   // 2^-16, relative position error of the FP32 SAT stage

struct XTri {        // obstacle triangle in the robot frame (FP32); 13 words = odd stride, bank-conflict free
  float v[9];
  float err;         // absolute position error bound of these 9 values
  float mabs;        // max |v|
  int tri;           // triangle index (leaf order) for the exact stage
  int pad;
};

#ifndef SFFG_CAND_CAP
#define SFFG_CAND_CAP 192
#endif
constexpr int kCandCap = SFFG_CAND_CAP;    // candidate triangles of one group of edge samples (swept-box traversal)

struct WarpScratch {
  int stack[kStackCap];
  int tri[kTriCap];
  int cand[kCandCap];
  XTri xt[32];
};

// what a CTA stages once: the robot records, the top cut and the first nodes of the (breadth-first) hierarchy
struct CtaShared {
  const RobotTri *rob;
  const float4 *top;     // 2 float4 per slot, E.n_top slots
  const float4 *nodes;   // 16 float4 per node, n_stage nodes
  int n_stage;
};
constexpr int kTopSlots = 32;

// ---------------------------------------------------------------------------------------------------------
// FP32 certificate SAT
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float min3f(float a, float b, float c) { return fminf(a, fminf(b, c)); }
__device__ __forceinline__ float max3f(float a, float b, float c) { return fmaxf(a, fmaxf(b, c)); }

// true when axis (ax,ay,az) separates the two triangles by more than the rounding bound
__device__ __forceinline__ bool axis_certifies(float ax, float ay, float az, const float *p, const float (*q)[3],
                                               float errpos) {
  float P1 = ax * p[0] + ay * p[1] + az * p[2];
  float P2 = ax * p[3] + ay * p[4] + az * p[5];
  float P3 = ax * p[6] + ay * p[7] + az * p[8];
  float Q1 = ax * q[0][0] + ay * q[0][1] + az * q[0][2];
  float Q2 = ax * q[1][0] + ay * q[1][1] + az * q[1][2];
  float Q3 = ax * q[2][0] + ay * q[2][1] + az * q[2][2];
  float gap = fmaxf(min3f(P1, P2, P3) - max3f(Q1, Q2, Q3), min3f(Q1, Q2, Q3) - max3f(P1, P2, P3));
  float bound = (fabsf(ax) + fabsf(ay) + fabsf(az)) * errpos;
  return gap > bound;
}

// axis with a host-precomputed (outward rounded) projection interval of the robot triangle
__device__ __forceinline__ bool axis_certifies_pre(const float *a, const float *p, float qlo, float qhi, float errpos) {
  float P1 = a[0] * p[0] + a[1] * p[1] + a[2] * p[2];
  float P2 = a[0] * p[3] + a[1] * p[4] + a[2] * p[5];
  float P3 = a[0] * p[6] + a[1] * p[7] + a[2] * p[8];
  float gap = fmaxf(min3f(P1, P2, P3) - qhi, qlo - max3f(P1, P2, P3));
  float bound = (fabsf(a[0]) + fabsf(a[1]) + fabsf(a[2])) * errpos;
  return gap > bound;
}

// Stage P1, lane-per-pair: the 8 most selective certificates -- AABB in the robot frame, the two face normals and the
// six in-plane edge normals (robot-side intervals precomputed on the host).
// Returns true when the pair is PROVEN disjoint; false = still open (stage P2 looks at it).
__device__ bool pair_quick_disjoint(const XTri &x, const RobotTri &rt) {
  const float *p = x.v;
  const float errpos = 2.0f * x.err + kEpsSat * fmaxf(x.mabs, rt.qmax);
  {
    float lo0 = min3f(p[0], p[3], p[6]) - errpos, hi0 = max3f(p[0], p[3], p[6]) + errpos;
    float lo1 = min3f(p[1], p[4], p[7]) - errpos, hi1 = max3f(p[1], p[4], p[7]) + errpos;
    float lo2 = min3f(p[2], p[5], p[8]) - errpos, hi2 = max3f(p[2], p[5], p[8]) + errpos;
    if (lo0 > rt.hi[0] || hi0 < rt.lo[0] || lo1 > rt.hi[1] || hi1 < rt.lo[1] || lo2 > rt.hi[2] || hi2 < rt.lo[2])
      return true;
  }
  if (axis_certifies_pre(rt.m, p, rt.m_lo, rt.m_hi, errpos)) return true;
  const float e0[3] = {p[3] - p[0], p[4] - p[1], p[5] - p[2]}, e1[3] = {p[6] - p[3], p[7] - p[4], p[8] - p[5]};
  const float n0 = e0[1] * e1[2] - e0[2] * e1[1], n1 = e0[2] * e1[0] - e0[0] * e1[2], n2 = e0[0] * e1[1] - e0[1] * e1[0];
  if (axis_certifies(n0, n1, n2, p, rt.q, errpos)) return true;
#pragma unroll
  for (int k = 0; k < 3; ++k)
    if (axis_certifies_pre(rt.h[k], p, rt.h_lo[k], rt.h_hi[k], errpos)) return true;
  // obstacle in-plane edge normals e_i x n: what separates a robot lying in the plane of a big triangle but beyond an edge
  const float e2[3] = {p[0] - p[6], p[1] - p[7], p[2] - p[8]};
  if (axis_certifies(e0[1] * n2 - e0[2] * n1, e0[2] * n0 - e0[0] * n2, e0[0] * n1 - e0[1] * n0, p, rt.q, errpos)) return true;
  if (axis_certifies(e1[1] * n2 - e1[2] * n1, e1[2] * n0 - e1[0] * n2, e1[0] * n1 - e1[1] * n0, p, rt.q, errpos)) return true;
  if (axis_certifies(e2[1] * n2 - e2[2] * n1, e2[2] * n0 - e2[0] * n2, e2[0] * n1 - e2[1] * n0, p, rt.q, errpos)) return true;
  return false;
}

// ---------------------------------------------------------------------------------------------------------
// FP32 contact certificate.  A pair the separating-axis stage could not clear is usually a deep contact; proving it
// in FP32 saves the FP64 stage.  Sufficient condition: an edge of one triangle pierces the interior of the other --
// its endpoints lie strictly on opposite sides of the other triangle's plane and the three signed volumes
// [(c_k - a) x (c_{k+1} - a)] . (b - a) share a sign -- with every quantity clear of a rigorous error bound
// (positions are off by at most errpos; a bilinear / trilinear form of vectors bounded by 2M then moves by less than
// 150 * errpos * M^2, allotted 256).  Lane j < 6 checks one (edge, triangle) combination; any lane suffices.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool edge_pierces(const float *a, const float *b, const float *c0, const float *c1, const float *c2,
                                             float bound) {
  const float u[3] = {c1[0] - c0[0], c1[1] - c0[1], c1[2] - c0[2]}, v[3] = {c2[0] - c0[0], c2[1] - c0[1], c2[2] - c0[2]};
  const float n[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
  const float da = n[0] * (a[0] - c0[0]) + n[1] * (a[1] - c0[1]) + n[2] * (a[2] - c0[2]);
  const float db = n[0] * (b[0] - c0[0]) + n[1] * (b[1] - c0[1]) + n[2] * (b[2] - c0[2]);
  if (!((da > bound && db < -bound) || (da < -bound && db > bound))) return false;
  const float d[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
  const float w0[3] = {c0[0] - a[0], c0[1] - a[1], c0[2] - a[2]}, w1[3] = {c1[0] - a[0], c1[1] - a[1], c1[2] - a[2]},
              w2[3] = {c2[0] - a[0], c2[1] - a[1], c2[2] - a[2]};
  const float v0 = (w0[1] * w1[2] - w0[2] * w1[1]) * d[0] + (w0[2] * w1[0] - w0[0] * w1[2]) * d[1] + (w0[0] * w1[1] - w0[1] * w1[0]) * d[2];
  const float v1 = (w1[1] * w2[2] - w1[2] * w2[1]) * d[0] + (w1[2] * w2[0] - w1[0] * w2[2]) * d[1] + (w1[0] * w2[1] - w1[1] * w2[0]) * d[2];
  const float v2 = (w2[1] * w0[2] - w2[2] * w0[1]) * d[0] + (w2[2] * w0[0] - w2[0] * w0[2]) * d[1] + (w2[0] * w0[1] - w2[1] * w0[0]) * d[2];
  return (v0 > bound && v1 > bound && v2 > bound) || (v0 < -bound && v1 < -bound && v2 < -bound);
}

// Stage P2, ten lanes per open pair, three pairs per pass (lanes 0..9, 10..19, 20..29; k = lane % 10):
//   pass a  lanes k < 9 evaluate the nine edge x edge axes e_i x f_j        -> pairs proven disjoint
//   pass b  (only for pairs pass a left open) lanes k < 6 evaluate the six edge-pierces-triangle contact certificates
__device__ __forceinline__ bool open_pair_axis(const XTri &x, const RobotTri &rt, int k) {
  const float *p = x.v;
  const float errpos = 2.0f * x.err + kEpsSat * fmaxf(x.mabs, rt.qmax);
  const int i = k / 3, j = k - 3 * i, i1 = i == 2 ? 0 : i + 1;
  const float e[3] = {p[3 * i1] - p[3 * i], p[3 * i1 + 1] - p[3 * i + 1], p[3 * i1 + 2] - p[3 * i + 2]};
  const float ax = e[1] * rt.f[j][2] - e[2] * rt.f[j][1], ay = e[2] * rt.f[j][0] - e[0] * rt.f[j][2],
              az = e[0] * rt.f[j][1] - e[1] * rt.f[j][0];
  return axis_certifies(ax, ay, az, p, rt.q, errpos);
}
__device__ __forceinline__ bool open_pair_pierce(const XTri &x, const RobotTri &rt, int c) {
  const float M = fmaxf(x.mabs, rt.qmax);
  const float bound = 256.0f * (2.0f * x.err + kEpsSat * M) * M * M;
  const int ed = c < 3 ? c : c - 3, ed1 = ed == 2 ? 0 : ed + 1;
  const bool obst_edge = c < 3;
  const float *a = obst_edge ? x.v + 3 * ed : rt.q[ed], *b = obst_edge ? x.v + 3 * ed1 : rt.q[ed1];
  const float *c0 = obst_edge ? rt.q[0] : x.v, *c1 = obst_edge ? rt.q[1] : x.v + 3, *c2 = obst_edge ? rt.q[2] : x.v + 6;
  return edge_pierces(a, b, c0, c1, c2, bound);
}

// ---------------------------------------------------------------------------------------------------------
// FP64 exact stage -- operation order of oracle/sff_oracle.c (make_xform, xform_point, orc_tri_contact).
// Explicit *_rn intrinsics keep ptxas from contracting a*b+c into an FMA, which the CPU oracle does not do.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double ddot(const double *a, const double *b) {
  return dadd(dadd(dmul(a[0], b[0]), dmul(a[1], b[1])), dmul(a[2], b[2]));
}
__device__ __forceinline__ void dcross(double *r, const double *a, const double *b) {
  r[0] = dsub(dmul(a[1], b[2]), dmul(a[2], b[1]));
  r[1] = dsub(dmul(a[2], b[0]), dmul(a[0], b[2]));
  r[2] = dsub(dmul(a[0], b[1]), dmul(a[1], b[0]));
}

// Point<T>::FillRotationMatrix (reference src/primitives.h:252-262), double.  The three sincos calls are spread over
// lanes 0..2 (every lane runs one call) and exchanged by shuffle; all lanes return the same matrix.
__device__ void rotation_f64_warp(double yaw, double pitch, double roll, double *m, int lane) {
  const int sel = lane % 3;
  const double ang = sel == 0 ? yaw : (sel == 1 ? pitch : roll);
  double s, c;
  sincos(ang, &s, &c);
  const double sy = __shfl_sync(kFull, s, 0), cy = __shfl_sync(kFull, c, 0);
  const double sp = __shfl_sync(kFull, s, 1), cp = __shfl_sync(kFull, c, 1);
  const double sr = __shfl_sync(kFull, s, 2), cr = __shfl_sync(kFull, c, 2);
  m[0] = dmul(cy, cp);
  m[1] = dsub(dmul(dmul(cy, sp), sr), dmul(sy, cr));
  m[2] = dadd(dmul(dmul(cy, sp), cr), dmul(sy, sr));
  m[3] = dmul(sy, cp);
  m[4] = dadd(dmul(dmul(sy, sp), sr), dmul(cy, cr));
  m[5] = dsub(dmul(dmul(sy, sp), cr), dmul(cy, sr));
  m[6] = -sp;
  m[7] = dmul(cp, sr);
  m[8] = dmul(cp, cr);
}

// all 32 lanes call this with identical arguments; lane k < 17 evaluates axis k.  Returns the pair verdict.
__device__ __noinline__ bool exact_pair_contact(const double *R2, const double *T2, const double *ot, const double *rq,
                                                int lane) {
  // mT = R2^T * (0 - T2); vertex = (R2^T row . p) * 1.0 + mT
  double u[3] = {dsub(0.0, T2[0]), dsub(0.0, T2[1]), dsub(0.0, T2[2])};
  double mT[3], P[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i) mT[i] = dadd(dadd(dmul(R2[0 + i], u[0]), dmul(R2[3 + i], u[1])), dmul(R2[6 + i], u[2]));
#pragma unroll
  for (int v = 0; v < 3; ++v)
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      double s = dadd(dadd(dmul(R2[0 + i], ot[3 * v]), dmul(R2[3 + i], ot[3 * v + 1])), dmul(R2[6 + i], ot[3 * v + 2]));
      P[v][i] = dadd(dmul(1.0, s), mT[i]);
    }
  // everything relative to P[0]
  double p[3][3], q[3][3], vec[8][3];   // vec: e1 e2 e3 f1 f2 f3 n1 m1
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    p[0][k] = dsub(P[0][k], P[0][k]);
    p[1][k] = dsub(P[1][k], P[0][k]);
    p[2][k] = dsub(P[2][k], P[0][k]);
    q[0][k] = dsub(rq[k], P[0][k]);
    q[1][k] = dsub(rq[3 + k], P[0][k]);
    q[2][k] = dsub(rq[6 + k], P[0][k]);
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    vec[0][k] = dsub(p[1][k], p[0][k]);
    vec[1][k] = dsub(p[2][k], p[1][k]);
    vec[2][k] = dsub(p[0][k], p[2][k]);
    vec[3][k] = dsub(q[1][k], q[0][k]);
    vec[4][k] = dsub(q[2][k], q[1][k]);
    vec[5][k] = dsub(q[0][k], q[2][k]);
  }
  dcross(vec[6], vec[0], vec[1]);
  dcross(vec[7], vec[3], vec[4]);
  // axis table: 0: n1, 1: m1, 2..10: e_i x f_j, 11..13: e_i x n1, 14..16: f_j x m1
  bool overlaps = true;
  if (lane < 17) {
    double ax[3];
    if (lane == 0) { ax[0] = vec[6][0]; ax[1] = vec[6][1]; ax[2] = vec[6][2]; }
    else if (lane == 1) { ax[0] = vec[7][0]; ax[1] = vec[7][1]; ax[2] = vec[7][2]; }
    else {
      int a, b;
      if (lane < 11) { a = (lane - 2) / 3; b = 3 + (lane - 2) % 3; }
      else if (lane < 14) { a = lane - 11; b = 6; }
      else { a = 3 + (lane - 14); b = 7; }
      dcross(ax, vec[a], vec[b]);
    }
    double P1 = ddot(ax, p[0]), P2 = ddot(ax, p[1]), P3 = ddot(ax, p[2]);
    double Q1 = ddot(ax, q[0]), Q2 = ddot(ax, q[1]), Q3 = ddot(ax, q[2]);
    double mx1 = fmax(P1, fmax(P2, P3)), mn1 = fmin(P1, fmin(P2, P3));
    double mx2 = fmax(Q1, fmax(Q2, Q3)), mn2 = fmin(Q1, fmin(Q2, Q3));
    if (mn1 > mx2) overlaps = false;
    if (mn2 > mx1) overlaps = false;
  }
  return __all_sync(kFull, overlaps);
}

// ---------------------------------------------------------------------------------------------------------
// pose formats.  Each lane keeps its own pose in the narrowest form that loses nothing; the exact stage fetches the
// double pose of lane `src` on demand (rare), so the FP32 fast path of an f32 batch never touches the FP64 pipe.
// ---------------------------------------------------------------------------------------------------------
enum { kFmtEulerF32 = 0, kFmtEulerF64 = 1, kFmtMatrixF64 = 2 };

template <int FMT> struct LanePose;

template <> struct LanePose<kFmtEulerF32> {
  float t[3], a[3];
  __device__ __forceinline__ void load(const void *base, long long i) {
    const float2 *p = reinterpret_cast<const float2 *>(base) + 3 * i;
    const float2 x = p[0], y = p[1], z = p[2];
    t[0] = x.x; t[1] = x.y; t[2] = y.x; a[0] = y.y; a[1] = z.x; a[2] = z.y;
  }
  __device__ __forceinline__ void clear() { t[0] = t[1] = t[2] = a[0] = a[1] = a[2] = 0.f; }
  __device__ __forceinline__ void split(float *hi, float *lo) const {
#pragma unroll
    for (int k = 0; k < 3; ++k) { hi[k] = t[k]; lo[k] = 0.f; }
  }
  __device__ __forceinline__ void rot32(float *R) const;
  __device__ __forceinline__ void exact(int src, double *R2, double *T2, int lane) const {
#pragma unroll
    for (int k = 0; k < 3; ++k) T2[k] = (double)__shfl_sync(kFull, t[k], src);
    rotation_f64_warp((double)__shfl_sync(kFull, a[0], src), (double)__shfl_sync(kFull, a[1], src),
                      (double)__shfl_sync(kFull, a[2], src), R2, lane);
  }
};

template <> struct LanePose<kFmtEulerF64> {
  double t[3], a[3];
  bool identity;   // edge samples in reference mode: R = I exactly, no trigonometry anywhere
  __device__ __forceinline__ void load(const void *base, long long i) {
    const double2 *p = reinterpret_cast<const double2 *>(base) + 3 * i;
    const double2 x = p[0], y = p[1], z = p[2];
    t[0] = x.x; t[1] = x.y; t[2] = y.x; a[0] = y.y; a[1] = z.x; a[2] = z.y;
    identity = false;
  }
  __device__ __forceinline__ void clear() { t[0] = t[1] = t[2] = a[0] = a[1] = a[2] = 0.0; identity = false; }
  __device__ __forceinline__ void split(float *hi, float *lo) const {
#pragma unroll
    for (int k = 0; k < 3; ++k) { hi[k] = (float)t[k]; lo[k] = (float)(t[k] - (double)hi[k]); }
  }
  __device__ __forceinline__ void rot32(float *R) const;
  __device__ __forceinline__ void exact(int src, double *R2, double *T2, int lane) const {
#pragma unroll
    for (int k = 0; k < 3; ++k) T2[k] = __shfl_sync(kFull, t[k], src);
    if (__shfl_sync(kFull, (int)identity, src)) {
      // FillRotationMatrix at zero angles: identity with m[2][0] = -sin(0) = -0.0
      R2[0] = 1.0; R2[1] = 0.0; R2[2] = 0.0; R2[3] = 0.0; R2[4] = 1.0; R2[5] = 0.0; R2[6] = -0.0; R2[7] = 0.0; R2[8] = 1.0;
    } else {
      rotation_f64_warp(__shfl_sync(kFull, a[0], src), __shfl_sync(kFull, a[1], src), __shfl_sync(kFull, a[2], src), R2, lane);
    }
  }
};

template <> struct LanePose<kFmtMatrixF64> {
  double t[3], r[9];   // RAPID_Collide's own arguments: R2 row-major, T2
  __device__ __forceinline__ void load(const void *base, long long i) {
    const double *p = reinterpret_cast<const double *>(base) + 12 * i;
#pragma unroll
    for (int k = 0; k < 9; ++k) r[k] = p[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) t[k] = p[9 + k];
  }
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int k = 0; k < 9; ++k) r[k] = 0.0;
    t[0] = t[1] = t[2] = 0.0;
  }
  __device__ __forceinline__ void split(float *hi, float *lo) const {
#pragma unroll
    for (int k = 0; k < 3; ++k) { hi[k] = (float)t[k]; lo[k] = (float)(t[k] - (double)hi[k]); }
  }
  __device__ __forceinline__ void rot32(float *R) const {
#pragma unroll
    for (int k = 0; k < 9; ++k) R[k] = (float)r[k];
  }
  __device__ __forceinline__ void exact(int src, double *R2, double *T2, int) const {
#pragma unroll
    for (int k = 0; k < 3; ++k) T2[k] = __shfl_sync(kFull, t[k], src);
#pragma unroll
    for (int k = 0; k < 9; ++k) R2[k] = __shfl_sync(kFull, r[k], src);
  }
};

// Point<T>::FillRotationMatrix in FP32 (certificate stage only)
__device__ __forceinline__ void rotation_f32(float yaw, float pitch, float roll, float *m) {
  float sy, cy, sp, cp, sr, cr;
  sincosf(yaw, &sy, &cy);
  sincosf(pitch, &sp, &cp);
  sincosf(roll, &sr, &cr);
  m[0] = cy * cp;
  m[1] = cy * sp * sr - sy * cr;
  m[2] = cy * sp * cr + sy * sr;
  m[3] = sy * cp;
  m[4] = sy * sp * sr + cy * cr;
  m[5] = sy * sp * cr - cy * sr;
  m[6] = -sp;
  m[7] = cp * sr;
  m[8] = cp * cr;
}
__device__ __forceinline__ void LanePose<kFmtEulerF32>::rot32(float *R) const { rotation_f32(a[0], a[1], a[2], R); }
__device__ __forceinline__ void LanePose<kFmtEulerF64>::rot32(float *R) const {
  if (identity) {
    R[0] = 1.f; R[1] = 0.f; R[2] = 0.f; R[3] = 0.f; R[4] = 1.f; R[5] = 0.f; R[6] = 0.f; R[7] = 0.f; R[8] = 1.f;
  } else {
    rotation_f32((float)a[0], (float)a[1], (float)a[2], R);
  }
}

// ---------------------------------------------------------------------------------------------------------
// per-pose warp traversal
// ---------------------------------------------------------------------------------------------------------
struct PoseU {          // warp-uniform copy of one pose
  float R[9];           // row-major, world <- robot
  float Thi[3], Tlo[3]; // T = Thi + Tlo (+ negligible)
};

struct Tally { unsigned long long past_root, box, pair, exact, steps, tri_passes, tris, exact_run, past_grid; };

struct BoxTest {        // per-pose constants of the oriented-box test
  float o[3], ra[3], rob_sz;
};

__device__ __forceinline__ BoxTest make_box_test(const EnvDev &E, const PoseU &P) {
  BoxTest bt;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    bt.o[k] = P.R[3 * k] * E.rob_c[0] + P.R[3 * k + 1] * E.rob_c[1] + P.R[3 * k + 2] * E.rob_c[2] + P.Tlo[k];
    bt.ra[k] = fabsf(P.R[3 * k]) * E.rob_h[0] + fabsf(P.R[3 * k + 1]) * E.rob_h[1] + fabsf(P.R[3 * k + 2]) * E.rob_h[2];
  }
  bt.rob_sz = 2.0f * E.rob_radius + fabsf(P.Tlo[0]) + fabsf(P.Tlo[1]) + fabsf(P.Tlo[2]);
  return bt;
}

// robot oriented box (centre T + R c, axes R, half extents h) against an AABB slot; conservative.  Straight-line code:
// lanes of one step almost never agree on an early exit, so all six axes are evaluated and combined without branches.
__device__ __forceinline__ bool slot_overlaps(const EnvDev &E, const PoseU &P, const BoxTest &bt, const float4 a, const float4 b) {
  const float tx = (a.x - P.Thi[0]) - bt.o[0], ty = (a.y - P.Thi[1]) - bt.o[1], tz = (a.z - P.Thi[2]) - bt.o[2];
  const float pad = kEpsBox * (fabsf(tx) + fabsf(ty) + fabsf(tz) + b.x + b.y + b.z + bt.rob_sz);
  const float s0 = P.R[0] * tx + P.R[3] * ty + P.R[6] * tz, s1 = P.R[1] * tx + P.R[4] * ty + P.R[7] * tz,
              s2 = P.R[2] * tx + P.R[5] * ty + P.R[8] * tz;
  const float r0 = fabsf(P.R[0]) * b.x + fabsf(P.R[3]) * b.y + fabsf(P.R[6]) * b.z,
              r1 = fabsf(P.R[1]) * b.x + fabsf(P.R[4]) * b.y + fabsf(P.R[7]) * b.z,
              r2 = fabsf(P.R[2]) * b.x + fabsf(P.R[5]) * b.y + fabsf(P.R[8]) * b.z;
  // the largest excess over the allowed distance on any axis; overlap iff none is positive
  const float ex = fmaxf(fmaxf(fabsf(tx) - (b.x + bt.ra[0]), fabsf(ty) - (b.y + bt.ra[1])), fabsf(tz) - (b.z + bt.ra[2]));
  const float eb = fmaxf(fmaxf(fabsf(s0) - (E.rob_h[0] + r0), fabsf(s1) - (E.rob_h[1] + r1)), fabsf(s2) - (E.rob_h[2] + r2));
  return !(fmaxf(ex, eb) > pad);
}

// state of the pair stages of one pose: the verdict so far and the pose's double-precision transform (built on first use)
// (the flags stay separate scalars: the transform arrays are handed to the non-inlined exact stage by address, and a struct
// holding both would drag the flags into local memory with them -- measured +2.4 % on the pose kernel)
struct PairCtx {
  double R2[9], T2[3];
};

// obstacle triangle t (leaf order) into the robot frame of pose P (what RAPID does: x = R^T (p - T)), tagged with the lane
// `tag` its pose came from; returns false when its box, padded by the position error bound, misses the robot's box
__device__ __forceinline__ bool transform_triangle(const EnvDev &E, const PoseU &P, int t, int tag, XTri &x) {
  const float4 v0 = __ldg(E.tris32 + 3 * (size_t)t), v1 = __ldg(E.tris32 + 3 * (size_t)t + 1),
               v2 = __ldg(E.tris32 + 3 * (size_t)t + 2);
  const float w[9] = {(v0.x - P.Thi[0]) - P.Tlo[0], (v0.y - P.Thi[1]) - P.Tlo[1], (v0.z - P.Thi[2]) - P.Tlo[2],
                      (v1.x - P.Thi[0]) - P.Tlo[0], (v1.y - P.Thi[1]) - P.Tlo[1], (v1.z - P.Thi[2]) - P.Tlo[2],
                      (v2.x - P.Thi[0]) - P.Tlo[0], (v2.y - P.Thi[1]) - P.Tlo[1], (v2.z - P.Thi[2]) - P.Tlo[2]};
  float mabs = 0.f;
#pragma unroll
  for (int v = 0; v < 3; ++v)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float val = P.R[c] * w[3 * v] + P.R[3 + c] * w[3 * v + 1] + P.R[6 + c] * w[3 * v + 2];
      x.v[3 * v + c] = val;
      mabs = fmaxf(mabs, fabsf(val));
    }
  x.err = v0.w;
  x.mabs = mabs;
  x.tri = t;
  x.pad = tag;
  const float padT = 2.0f * x.err + kEpsSat * fmaxf(mabs, E.rob_radius);
  bool keep = true;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float lo = min3f(x.v[c], x.v[3 + c], x.v[6 + c]) - padT, hi = max3f(x.v[c], x.v[3 + c], x.v[6 + c]) + padT;
    const float rc = c == 0 ? E.rob_c[0] : (c == 1 ? E.rob_c[1] : E.rob_c[2]);
    const float rh = c == 0 ? E.rob_h[0] : (c == 1 ? E.rob_h[1] : E.rob_h[2]);
    if (lo > rc + rh || hi < rc - rh) keep = false;
  }
  return keep;
}

// Triangle stage of one pose over the candidate triangles ws.tri[0, ntri): transform into the robot frame, cull, pair stages
// P1 (lane-per-pair), P2 (three pairs per cooperative pass) and the FP64 exact stage; stops at the first confirmed contact.
template <int FMT, bool COUNT>
__device__ __forceinline__ void triangle_stage(const EnvDev &E, WarpScratch &ws, const RobotTri *srob, const PoseU &P,
                                               const LanePose<FMT> &lp, int src, int lane, int ntri, bool &hit, bool &have_R2,
                                               PairCtx &pc, Tally &tally) {
  const unsigned lt = (1u << lane) - 1u;
  const int grp10 = lane / 10, k10 = lane - 10 * grp10;
  const float inv_n_robot = 1.0f / (float)E.n_robot;
  for (int base = 0; base < ntri && !hit; base += 32) {
    const int cnt = (ntri - base) < 32 ? (ntri - base) : 32;
    bool keep = false;
    XTri x;
    if (lane < cnt) keep = transform_triangle(E, P, ws.tri[base + lane], 0, x);
    const unsigned km = __ballot_sync(kFull, keep);
    const int nx = __popc(km);
    if (COUNT) { tally.tri_passes += 1; tally.tris += cnt; }
    if (keep) ws.xt[__popc(km & lt)] = x;
    __syncwarp();
    const int npairs = nx * E.n_robot;
    for (int pb = 0; pb < npairs && !hit; pb += 32) {
      const int pidx = pb + lane;
      bool undecided = false;
      // (triangle, robot triangle) of this lane's pair; the quotient by float reciprocal is exact for pidx < 2^20
      const int xi_l = (int)(((float)pidx + 0.5f) * inv_n_robot), r_l = pidx - xi_l * E.n_robot;
      if (pidx < npairs) undecided = !pair_quick_disjoint(ws.xt[xi_l], srob[r_l]);
      unsigned um = __ballot_sync(kFull, undecided);
      if (COUNT) {
        const int np = (npairs - pb) < 32 ? (npairs - pb) : 32;
        tally.pair += np;
        tally.exact += __popc(um);   // pairs the lane-per-pair stage P1 left open
      }
      while (um && !hit) {
        const int l0 = __ffs(um) - 1;
        um &= um - 1;
        const int l1 = um ? __ffs(um) - 1 : -1;
        um &= um - 1;   // (0 stays 0)
        const int l2 = um ? __ffs(um) - 1 : -1;
        um &= um - 1;
        const int lmine = grp10 == 0 ? l0 : (grp10 == 1 ? l1 : (grp10 == 2 ? l2 : -1));
        const int xim = __shfl_sync(kFull, xi_l, lmine & 31), rm = __shfl_sync(kFull, r_l, lmine & 31);
        bool sep = false;
        if (lmine >= 0 && k10 < 9) sep = open_pair_axis(ws.xt[xim], srob[rm], k10);
        const unsigned bs = __ballot_sync(kFull, sep);
        unsigned open_g = 0;
        if (!(bs & 0x3ffu)) open_g |= 1u;
        if (l1 >= 0 && !(bs & (0x3ffu << 10))) open_g |= 2u;
        if (l2 >= 0 && !(bs & (0x3ffu << 20))) open_g |= 4u;
        if (open_g == 0) continue;
        bool con = false;
        if (lmine >= 0 && k10 < 6 && ((open_g >> grp10) & 1u)) con = open_pair_pierce(ws.xt[xim], srob[rm], k10);
        if (__any_sync(kFull, con)) {
          hit = true;
          break;
        }
        while (open_g) {
          const int g = __ffs(open_g) - 1;
          open_g &= open_g - 1;
          const int lg = g == 0 ? l0 : (g == 1 ? l1 : l2);
          const int xg = __shfl_sync(kFull, xi_l, lg), rg = __shfl_sync(kFull, r_l, lg);
          if (COUNT) tally.exact_run += 1;
          if (!have_R2) {
            lp.exact(src, pc.R2, pc.T2, lane);
            have_R2 = true;
          }
          const int t = ws.xt[xg].tri;
          if (exact_pair_contact(pc.R2, pc.T2, E.tris64 + 9 * (size_t)t, E.robot64 + 9 * (size_t)rg, lane)) {
            hit = true;
            break;
          }
        }
      }
    }
    __syncwarp();
  }
}

template <int FMT, bool COUNT>
__device__ bool warp_pose_hit(const EnvDev &E, WarpScratch &ws, const CtaShared &cs, const PoseU &P,
                              const LanePose<FMT> &lp, int src, int lane, Tally &tally) {
  const RobotTri *srob = cs.rob;
  const BoxTest bt = make_box_test(E, P);
  const unsigned lt = (1u << lane) - 1u;

  int sp = 0, ntri = 0;
  bool hit = false, have_R2 = false;
  PairCtx pc;
  if (COUNT) tally.past_root += 1;

  bool first = true;   // step 0 tests the precomputed <=32-box cut of the top of the hierarchy with all lanes
  while (true) {
    if (first || (sp > 0 && ntri <= kTriCap - 32)) {
      bool ov = false;
      int child = kEmptyChild;
      bool active;
      if (first) {
        active = lane < E.n_top;
        if (active) {
          const float4 a = cs.top[lane], b = cs.top[kTopSlots + lane];
          child = __float_as_int(a.w);
          ov = slot_overlaps(E, P, bt, a, b);
        }
        first = false;
      } else {
        // 4 nodes per step while there is room; near the cap fall back to strict depth-first (1 node per step grows the
        // stack by at most 7 per level), which keeps any hierarchy of depth <= 45 inside the stack
        const int take = sp > kStackCap - 64 ? 1 : (sp < 4 ? sp : 4);
        const int grp = lane >> 3;
        active = grp < take;
        const int node = active ? ws.stack[sp - 1 - grp] : 0;
        __syncwarp();
        sp -= take;
        if (active) {
          float4 a, b;
          if (node < cs.n_stage) {   // (nodes of one step are popped together: mostly one side of this branch)
            const int si = node * kWide + (lane & 7);
            a = cs.nodes[si];                             // staged as two arrays (all `a` halves, then all `b` halves):
            b = cs.nodes[cs.n_stage * kWide + si];        // 8 lanes read 128 contiguous bytes, no bank conflicts
          } else {
            const float4 *gn = E.slots + ((size_t)node * kWide + (lane & 7)) * 2;
            a = __ldg(gn);
            b = __ldg(gn + 1);
          }
          child = __float_as_int(a.w);
          if (child != kEmptyChild) ov = slot_overlaps(E, P, bt, a, b);
        }
      }
      const unsigned m_int = __ballot_sync(kFull, ov && child >= 0);
      const unsigned m_leaf = __ballot_sync(kFull, ov && child < 0);
      if (COUNT) { tally.box += __popc(__ballot_sync(kFull, active && child != kEmptyChild)); tally.steps += 1; }
      if (ov && child >= 0) {
        const int pos = sp + __popc(m_int & lt);
        if (pos < kStackCap) ws.stack[pos] = child;
      }
      if (ov && child < 0) ws.tri[ntri + __popc(m_leaf & lt)] = ~child;
      sp += __popc(m_int);
      ntri += __popc(m_leaf);
      if (sp > kStackCap) {     // never silently drop work: flag the launch as failed
        if (lane == 0) *reinterpret_cast<volatile int *>(E.status) = 1;
        sp = kStackCap;
      }
      __syncwarp();
      if (sp > 0 && ntri < kTriFlush) continue;
    }
    if (ntri == 0) {
      if (sp == 0) break;
      continue;
    }
    // ---------------- triangle stage ----------------
    triangle_stage<FMT, COUNT>(E, ws, srob, P, lp, src, lane, ntri, hit, have_R2, pc, tally);
    if (hit) break;
    ntri = 0;
    if (sp == 0) break;
  }
  return hit;
}

// ---------------------------------------------------------------------------------------------------------
// Edge samples in reference mode all carry the identity rotation (src/problemStruct.h:157-163), so the robot's box at a
// sample is the axis-aligned box T + [rob_c - rob_h, rob_c + rob_h] and the boxes of the surviving samples of one
// 32-sample group lie in one slightly larger axis-aligned box.  ONE traversal with that swept box collects every triangle
// any of the samples could touch; each sample then runs only the triangle stage over that list.  (The per-sample
// traversal was 63 % of the edge kernel's instructions.)  The list is a superset of what the per-sample traversal would
// deliver and every stage after it is unchanged, so the verdicts are the same.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int f2ord(float f) {   // order-preserving float -> int
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

// every leaf triangle whose slot box meets [lo, hi] -> ws.cand; returns their number, or -1 when the list would overflow
template <bool COUNT>
__device__ int collect_box_candidates(const EnvDev &E, WarpScratch &ws, const CtaShared &cs, const float *lo, const float *hi, int lane,
                                      Tally &tally) {
  const unsigned lt = (1u << lane) - 1u;
  int sp = 0, ncand = 0;
  bool first = true;
  while (first || sp > 0) {
    bool ov = false;
    int child = kEmptyChild;
    bool active;
    float4 a, b;
    if (first) {
      active = lane < E.n_top;
      if (active) {
        a = cs.top[lane];
        b = cs.top[kTopSlots + lane];
      }
      first = false;
    } else {
      const int take = sp > kStackCap - 64 ? 1 : (sp < 4 ? sp : 4);
      const int grp = lane >> 3;
      active = grp < take;
      const int node = active ? ws.stack[sp - 1 - grp] : 0;
      __syncwarp();
      sp -= take;
      if (active) {
        if (node < cs.n_stage) {
          const int si = node * kWide + (lane & 7);
          a = cs.nodes[si];
          b = cs.nodes[cs.n_stage * kWide + si];
        } else {
          const float4 *gn = E.slots + ((size_t)node * kWide + (lane & 7)) * 2;
          a = __ldg(gn);
          b = __ldg(gn + 1);
        }
      }
    }
    if (active) {
      child = __float_as_int(a.w);
      // slot box [c - h, c + h] against [lo, hi], directed rounding keeps the test conservative
      if (child != kEmptyChild)
        ov = __fadd_ru(a.x, b.x) >= lo[0] && __fsub_rd(a.x, b.x) <= hi[0] && __fadd_ru(a.y, b.y) >= lo[1] && __fsub_rd(a.y, b.y) <= hi[1] &&
             __fadd_ru(a.z, b.z) >= lo[2] && __fsub_rd(a.z, b.z) <= hi[2];
    }
    const unsigned m_int = __ballot_sync(kFull, ov && child >= 0);
    const unsigned m_leaf = __ballot_sync(kFull, ov && child < 0);
    if (COUNT) { tally.box += __popc(__ballot_sync(kFull, active && child != kEmptyChild)); tally.steps += 1; }
    if (ncand + __popc(m_leaf) > kCandCap) return -1;
    if (ov && child >= 0) {
      const int pos = sp + __popc(m_int & lt);
      if (pos < kStackCap) ws.stack[pos] = child;
    }
    if (ov && child < 0) ws.cand[ncand + __popc(m_leaf & lt)] = ~child;
    sp += __popc(m_int);
    ncand += __popc(m_leaf);
    if (sp > kStackCap) return -1;   // (the caller falls back to the per-sample traversal, which reports overflows)
    __syncwarp();
  }
  return ncand;
}

// one sample against the shared candidate list
template <int FMT, bool COUNT>
__device__ bool pose_hits_candidates(const EnvDev &E, WarpScratch &ws, const CtaShared &cs, const PoseU &P, const LanePose<FMT> &lp,
                                     int src, int lane, int ncand, Tally &tally) {
  bool hit = false, have_R2 = false;
  PairCtx pc;
  if (COUNT) tally.past_root += 1;
  for (int base = 0; base < ncand && !hit; base += 32) {
    const int cnt = (ncand - base) < 32 ? (ncand - base) : 32;
    if (lane < cnt) ws.tri[lane] = ws.cand[base + lane];
    __syncwarp();
    triangle_stage<FMT, COUNT>(E, ws, cs.rob, P, lp, src, lane, cnt, hit, have_R2, pc, tally);
  }
  return hit;
}

// O(1) free-space test: the cell of the clearance grid that holds the robot origin.  A clear bit proves that no obstacle
// triangle comes within the robot's bounding radius of any point of the cell (the build inflates the reach by the cell's
// half diagonal plus rounding margins), so the pose is free whatever its orientation.
__device__ __forceinline__ bool clearance_says_free(const EnvDev &E, float tx, float ty, float tz) {
  const float fx = (tx - E.grid_o[0]) * E.grid_inv_h, fy = (ty - E.grid_o[1]) * E.grid_inv_h, fz = (tz - E.grid_o[2]) * E.grid_inv_h;
  if (!(fx >= 0.f && fy >= 0.f && fz >= 0.f && fx < (float)E.grid_n[0] && fy < (float)E.grid_n[1] && fz < (float)E.grid_n[2]))
    return false;
  const unsigned cell = ((unsigned)fz * (unsigned)E.grid_n[1] + (unsigned)fy) * (unsigned)E.grid_n[0] + (unsigned)fx;
  return ((__ldg(E.clear_bits + (cell >> 5)) >> (cell & 31)) & 1u) == 0u;
}

// lane-per-pose cull: robot bounding sphere (about the robot origin) against the obstacle AABB
__device__ __forceinline__ bool sphere_hits_root(const EnvDev &E, float tx, float ty, float tz, float tlo_mag) {
  const float dx = fmaxf(fabsf(tx - E.root_c[0]) - E.root_h[0], 0.f);
  const float dy = fmaxf(fabsf(ty - E.root_c[1]) - E.root_h[1], 0.f);
  const float dz = fmaxf(fabsf(tz - E.root_c[2]) - E.root_h[2], 0.f);
  const float pad = kEpsBox * (fabsf(tx) + fabsf(ty) + fabsf(tz) + fabsf(E.root_c[0]) + fabsf(E.root_c[1]) +
                               fabsf(E.root_c[2]) + E.root_h[0] + E.root_h[1] + E.root_h[2] + E.rob_radius) + tlo_mag;
  const float r = E.rob_radius + pad;
  return dx * dx + dy * dy + dz * dz <= r * r * 1.000001f;
}

// robot records, top cut and the first `n_stage` nodes of the hierarchy -> shared memory (once per CTA)
__device__ __forceinline__ CtaShared stage_cta(const EnvDev &E, unsigned char *smem, int n_stage) {
  RobotTri *srob = reinterpret_cast<RobotTri *>(smem);
  float4 *stop = reinterpret_cast<float4 *>(smem + (size_t)E.n_robot * sizeof(RobotTri));
  float4 *snodes = stop + 2 * kTopSlots;
  {
    const int words = E.n_robot * (int)(sizeof(RobotTri) / 4);
    const float *g = reinterpret_cast<const float *>(E.robot);
    float *d = reinterpret_cast<float *>(srob);
    for (int i = threadIdx.x; i < words; i += blockDim.x) d[i] = __ldg(g + i);
  }
  // slot = 2 float4 (a, b) in global memory; shared memory keeps all `a` halves first, then all `b` halves
  for (int i = threadIdx.x; i < 2 * E.n_top; i += blockDim.x) stop[(i & 1) * kTopSlots + (i >> 1)] = __ldg(E.top + i);
  const int nv = n_stage * kWide * 2;
  for (int i = threadIdx.x; i < nv; i += blockDim.x) snodes[(i & 1) * n_stage * kWide + (i >> 1)] = __ldg(E.slots + i);
  __syncthreads();
  CtaShared cs;
  cs.rob = srob;
  cs.top = stop;
  cs.nodes = snodes;
  cs.n_stage = n_stage;
  return cs;
}
__device__ __forceinline__ WarpScratch &warp_scratch(const EnvDev &E, unsigned char *smem, int warp) {
  unsigned char *base = smem + (size_t)E.n_robot * sizeof(RobotTri) + (size_t)2 * kTopSlots * sizeof(float4) +
                        (size_t)E.n_stage_max * kWide * 2 * sizeof(float4);
  return reinterpret_cast<WarpScratch *>(base)[warp];
}

__device__ __forceinline__ void flush_tally(const EnvDev &E, const Tally &t, unsigned long long poses, int lane) {
  if (lane == 0 && E.counters) {
    atomicAdd(E.counters + 0, poses);
    atomicAdd(E.counters + 1, t.past_root);
    atomicAdd(E.counters + 2, t.box);
    atomicAdd(E.counters + 3, t.pair);
    atomicAdd(E.counters + 4, t.exact);
    atomicAdd(E.counters + 8, t.exact_run);
    atomicAdd(E.counters + 9, t.past_grid);
    atomicAdd(E.counters + 5, t.steps);
    atomicAdd(E.counters + 6, t.tri_passes);
    atomicAdd(E.counters + 7, t.tris);
  }
}

// phase A (lane-per-pose cull + rotation) then phase B over the surviving lanes; returns the mask of colliding lanes
// (stops at the first hit when `first_only`, which is what an edge needs)
template <int FMT, bool COUNT, bool SWEPT_CAPABLE = false>
__device__ __forceinline__ unsigned check_32_poses(const EnvDev &E, WarpScratch &ws, const CtaShared &cs, bool valid,
                                                   const LanePose<FMT> &lp, int lane, bool first_only, Tally &tally,
                                                   bool swept = false, bool swept_identity = true) {
  float thi[3], tlo[3], R[9];
  lp.split(thi, tlo);
  bool alive = valid && E.n_obst > 0 &&
               sphere_hits_root(E, thi[0], thi[1], thi[2], fabsf(tlo[0]) + fabsf(tlo[1]) + fabsf(tlo[2]));
  if (alive && E.grid_n[0] > 0) alive = !clearance_says_free(E, thi[0], thi[1], thi[2]);
  if (COUNT) tally.past_grid += __popc(__ballot_sync(kFull, alive));
  if (alive) lp.rot32(R);
  unsigned todo = __ballot_sync(kFull, alive);
  unsigned hitmask = 0;
  if (SWEPT_CAPABLE && swept && __popc(todo) >= 2) {
    // one traversal with the box that holds the robot at every surviving sample, outward rounded: with the identity
    // rotation (reference mode) the robot's own axis-aligned box, with interpolated angles the cube around its bounding
    // sphere (any rotation keeps the robot inside it)
    int blo[3], bhi[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float c = __fadd_rd(thi[k], tlo[k]), cu = __fadd_ru(thi[k], tlo[k]);
      const float pad = kEpsBox * (fabsf(cu) + E.rob_radius);
      const float rl = swept_identity ? __fsub_rd(E.rob_c[k], E.rob_h[k]) : -E.rob_radius;
      const float rh = swept_identity ? __fadd_ru(E.rob_c[k], E.rob_h[k]) : E.rob_radius;
      const float l = __fsub_rd(__fadd_rd(c, rl), pad);
      const float h = __fadd_ru(__fadd_ru(cu, rh), pad);
      blo[k] = __reduce_min_sync(kFull, alive ? f2ord(l) : 0x7fffffff);
      bhi[k] = __reduce_max_sync(kFull, alive ? f2ord(h) : (int)0x80000000);
    }
    const float lo[3] = {ord2f(blo[0]), ord2f(blo[1]), ord2f(blo[2])}, hi[3] = {ord2f(bhi[0]), ord2f(bhi[1]), ord2f(bhi[2])};
    const int ncand = collect_box_candidates<COUNT>(E, ws, cs, lo, hi, lane, tally);
    if (ncand == 0) return 0;
    if (ncand > 0) {
      while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        PoseU P;
#pragma unroll
        for (int k = 0; k < 9; ++k) P.R[k] = __shfl_sync(kFull, R[k], src);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          P.Thi[k] = __shfl_sync(kFull, thi[k], src);
          P.Tlo[k] = FMT == kFmtEulerF32 ? 0.f : __shfl_sync(kFull, tlo[k], src);
        }
        if (pose_hits_candidates<FMT, COUNT>(E, ws, cs, P, lp, src, lane, ncand, tally)) {
          hitmask |= 1u << src;
          if (first_only) break;
        }
      }
      return hitmask;
    }
    // (list overflow: fall through to the per-sample traversal)
  }
  while (todo) {
    const int src = __ffs(todo) - 1;
    todo &= todo - 1;
    PoseU P;
#pragma unroll
    for (int k = 0; k < 9; ++k) P.R[k] = __shfl_sync(kFull, R[k], src);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      P.Thi[k] = __shfl_sync(kFull, thi[k], src);
      P.Tlo[k] = FMT == kFmtEulerF32 ? 0.f : __shfl_sync(kFull, tlo[k], src);
    }
    if (warp_pose_hit<FMT, COUNT>(E, ws, cs, P, lp, src, lane, tally)) {
      hitmask |= 1u << src;
      if (first_only) break;
    }
  }
  return hitmask;
}


// spins (one thread) until every rank has published an epoch >= `epoch` in the local flag words; bounded by a 10 s timeout
__device__ __forceinline__ void wait_flags(const FlagSet &f, unsigned epoch, int *status) {
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (int r = 0; r < f.n; ++r) {
    const unsigned *mine = f.p[f.me] + r;
    for (;;) {
      unsigned v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
      if ((int)(v - epoch) >= 0) break;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 10000000000ull) {   // a peer died or never launched -- flag it, never hang the GPU
        if (status) atomicExch(status, 3);
        else __trap();                  // no status word to raise (barrier without an environment): fail the launch loudly
        return;
      }
      __nanosleep(100);
    }
  }
}

template <int FMT, bool COUNT>
__global__ void __launch_bounds__(kThreads, SFFG_MIN_BLOCKS) collide_poses_kernel(EnvDev E, const void *poses, long long n,
                                                                                  OutSet outs, GatherSync gs, int chunk) {
  extern __shared__ __align__(16) unsigned char smem[];
  if (gs.flags.n > 0 && gs.wait_epoch != 0 && threadIdx.x == 0) wait_flags(gs.flags, gs.wait_epoch, E.status);
  const CtaShared cs = stage_cta(E, smem, E.n_stage_max);   // (its __syncthreads also releases the CTA from the wait above)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  WarpScratch &ws = warp_scratch(E, smem, warp);
  // a work unit is `chunk` (1..32) consecutive poses: 32 for large batches (full lanes in phase A), fewer when the
  // batch is too small to give every resident warp a unit (planner-sized calls are latency-, not throughput-bound)
  const long long nchunks = (n + chunk - 1) / chunk;
  Tally tally = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  unsigned long long nposes = 0;
  while (true) {
    unsigned c = 0;
    if (lane == 0) c = atomicAdd(E.work_counter, 1u) - E.work_base;
    c = __shfl_sync(kFull, c, 0);
    if ((long long)c >= nchunks) break;
    const long long i = (long long)c * chunk + lane;
    const bool mine = lane < chunk && i < n;
    LanePose<FMT> lp;
    if (mine) lp.load(poses, i);
    else lp.clear();
    if (COUNT) nposes += mine ? 1 : 0;
    const unsigned hitmask = check_32_poses<FMT, COUNT>(E, ws, cs, mine, lp, lane, false, tally);
    if (outs.n == 1) {
      if (mine) outs.p[0][i] = (uint8_t)((hitmask >> lane) & 1u);
    } else if (chunk == 32 && (long long)c * 32 + 32 <= n) {
      // peer-store gather: the 32 verdict bytes of this unit are 8 words; lane = 8 * destination + word, so one store
      // instruction serves four destinations with one full 32-byte segment each (local HBM or a peer over NVLink)
      const unsigned nib = (hitmask >> (4 * (lane & 7))) & 0xFu;
      const unsigned word = (nib & 1u) | ((nib & 2u) << 7) | ((nib & 4u) << 14) | ((nib & 8u) << 21);
      for (int r0 = 0; r0 < outs.n; r0 += 4) {
        const int r = r0 + (lane >> 3);
        if (r < outs.n) reinterpret_cast<unsigned *>(outs.p[r] + (long long)c * 32)[lane & 7] = word;
      }
    } else if (mine) {
      const uint8_t v = (uint8_t)((hitmask >> lane) & 1u);
      for (int r = 0; r < outs.n; ++r) outs.p[r][i] = v;
    }
  }
  if (COUNT) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) nposes += __shfl_xor_sync(kFull, nposes, s);
    flush_tally(E, tally, nposes, lane);
  }
  if (gs.flags.n > 0) {
    // completion signal of the fused gather: the last CTA to get here publishes the epoch to every rank
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence_system();
      if (atomicAdd(gs.done_counter, 1u) == gridDim.x - 1) {
        *gs.done_counter = 0;
        __threadfence_system();
        for (int r = 0; r < gs.flags.n; ++r)
          asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(gs.flags.p[r] + gs.flags.me), "r"(gs.signal_epoch) : "memory");
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// edges: Solver<T,R>::isPathFree (reference src/problemStruct.h:154-168), one warp per edge
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double wrap_pi(double a) {
  const double pi = 3.14159265358979323846;
  if (a < -pi) return dadd(a, dmul(2.0, pi));
  else if (a >= pi) return dsub(a, dmul(2.0, pi));
  return a;
}

// `split` (power of two) warps share one edge: warp j of an edge takes the j-th contiguous 1/split of the sample indices
// (neighbouring samples share one swept box), so a planner-sized batch still occupies the whole GPU.
// With split > 1 the per-edge minimum colliding index is combined with atomicMin in `fh` (pre-set to kNoHit) and warps
// stop as soon as an earlier hit than anything they could still find is published; finalize_edges_kernel then writes
// the outputs.  split == 1 writes them directly.
constexpr int kNoHit = 0x7f7f7f7f;   // what cudaMemsetAsync(..., 0x7f, ...) produces
// work units (edge, range) a planner-sized batch is cut into per resident warp.  Measured with the swept-box traversal
// (profiles/r02_edges_swept.md): 1 beats 2 / 4 / 8 for the planner's short edges (2 048 edges of length 4: 41.6 us per call
// against 44.9 / 50.9 / 69.6), long edges lose a little (length 12: 68 against 62 us at 4).
#ifndef SFFG_EDGE_UNITS_PER_WARP
#define SFFG_EDGE_UNITS_PER_WARP 1
#endif

#ifndef SFFG_EDGE_MIN_BLOCKS
#define SFFG_EDGE_MIN_BLOCKS SFFG_MIN_BLOCKS
#endif
template <bool COUNT>
__global__ void __launch_bounds__(kThreads, SFFG_EDGE_MIN_BLOCKS) check_edges_kernel(EnvDev E, const double *starts,
                                                                                const double *ends, long long m, double sample,
                                                                                int rot_mode, uint8_t *free_out,
                                                                                int32_t *first_hit, int split, int *fh) {
  extern __shared__ __align__(16) unsigned char smem[];
  const CtaShared cs = stage_cta(E, smem, E.n_stage_max);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  WarpScratch &ws = warp_scratch(E, smem, warp);
  Tally tally = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  unsigned long long nposes = 0;
  const long long units = m * split;
  while (true) {
    unsigned u = 0;
    if (lane == 0) u = atomicAdd(E.work_counter, 1u) - E.work_base;
    u = __shfl_sync(kFull, u, 0);
    if ((long long)u >= units) break;
    const unsigned eidx = u / (unsigned)split;
    const int part = (int)(u - eidx * (unsigned)split);
    double s[6], f[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      s[k] = __ldg(starts + 6 * (size_t)eidx + k);
      f[k] = __ldg(ends + 6 * (size_t)eidx + k);
    }
    // Point<T>::distance, src/primitives.h:224-235
    double sum = 0.0, adir[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double d = dsub(s[k], f[k]);
      sum = dadd(sum, dmul(d, d));
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      adir[k] = wrap_pi(dsub(f[3 + k], s[3 + k]));
      sum = dadd(sum, dmul(adir[k], adir[k]));
    }
    const double total = sqrt(sum);
    const double parts = __ddiv_rn(total, sample);
    const double dir[3] = {dsub(f[0], s[0]), dsub(f[1], s[1]), dsub(f[2], s[2])};
    // indices 1 .. S with (double)index < parts
    long long S = 0;
    if (parts > 1.0) {
      const double cl = ceil(parts);
      S = (cl > 2.0e9 ? 2000000000LL : (long long)cl) - 1;
    }
    int hit_index = 0;
    // warp `part` of the edge takes a CONTIGUOUS range of the sample indices (neighbouring samples share one swept box)
    const long long range_len = (S + split - 1) / split;
    const long long first_idx = 1 + (long long)part * range_len;
    const long long last_idx = first_idx + range_len - 1 < S ? first_idx + range_len - 1 : S;
    for (long long base = first_idx; base <= last_idx && hit_index == 0; base += 32) {
      if (split > 1 && *reinterpret_cast<volatile int *>(fh + eidx) < base) break;   // an earlier hit is already known
      const long long idx = base + lane;
      const bool valid = idx <= last_idx;
      const double di = (double)idx;
      LanePose<kFmtEulerF64> lp;
      lp.clear();
      lp.identity = rot_mode != SFFG_ROT_INTERPOLATE;
#pragma unroll
      for (int k = 0; k < 3; ++k) lp.t[k] = dadd(s[k], __ddiv_rn(dmul(di, dir[k]), parts));
      if (rot_mode == SFFG_ROT_INTERPOLATE) {
#pragma unroll
        for (int k = 0; k < 3; ++k) lp.a[k] = dadd(s[3 + k], __ddiv_rn(dmul(di, adir[k]), parts));
      }
      if (COUNT) nposes += valid ? 1 : 0;
      const unsigned hm = check_32_poses<kFmtEulerF64, COUNT, true>(E, ws, cs, valid, lp, lane, true, tally, true,
                                                                    rot_mode != SFFG_ROT_INTERPOLATE);
      if (hm) hit_index = (int)(base + (long long)(__ffs(hm) - 1));
    }
    if (lane == 0) {
      if (split > 1) {
        if (hit_index) atomicMin(fh + eidx, hit_index);
      } else {
        free_out[eidx] = hit_index == 0 ? 1 : 0;
        if (first_hit) first_hit[eidx] = hit_index;
      }
    }
  }
  if (COUNT) {
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) nposes += __shfl_xor_sync(kFull, nposes, sft);
    flush_tally(E, tally, nposes, lane);
  }
}

__global__ void finalize_edges_kernel(const int *fh, long long m, uint8_t *free_out, int32_t *first_hit) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const int v = fh[i];
  free_out[i] = v == kNoHit ? 1 : 0;
  if (first_hit) first_hit[i] = v == kNoHit ? 0 : v;
}

// ---------------------------------------------------------------------------------------------------------
// clearance grid build: one warp per obstacle triangle marks the cells whose centre lies within `reach` of the triangle
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float point_tri_dist2(const float *p, const float *a, const float *b, const float *c) {
  // closest point on a triangle (Voronoi-region walk), squared distance
  const float ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, ac[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
  const float ap[3] = {p[0] - a[0], p[1] - a[1], p[2] - a[2]};
  const float d1 = ab[0] * ap[0] + ab[1] * ap[1] + ab[2] * ap[2], d2 = ac[0] * ap[0] + ac[1] * ap[1] + ac[2] * ap[2];
  float q[3];
  if (d1 <= 0.f && d2 <= 0.f) { q[0] = a[0]; q[1] = a[1]; q[2] = a[2]; }
  else {
    const float bp[3] = {p[0] - b[0], p[1] - b[1], p[2] - b[2]};
    const float d3 = ab[0] * bp[0] + ab[1] * bp[1] + ab[2] * bp[2], d4 = ac[0] * bp[0] + ac[1] * bp[1] + ac[2] * bp[2];
    if (d3 >= 0.f && d4 <= d3) { q[0] = b[0]; q[1] = b[1]; q[2] = b[2]; }
    else {
      const float vc = d1 * d4 - d3 * d2;
      if (vc <= 0.f && d1 >= 0.f && d3 <= 0.f) {
        const float v = d1 / (d1 - d3);
        q[0] = a[0] + v * ab[0]; q[1] = a[1] + v * ab[1]; q[2] = a[2] + v * ab[2];
      } else {
        const float cp[3] = {p[0] - c[0], p[1] - c[1], p[2] - c[2]};
        const float d5 = ab[0] * cp[0] + ab[1] * cp[1] + ab[2] * cp[2], d6 = ac[0] * cp[0] + ac[1] * cp[1] + ac[2] * cp[2];
        if (d6 >= 0.f && d5 <= d6) { q[0] = c[0]; q[1] = c[1]; q[2] = c[2]; }
        else {
          const float vb = d5 * d2 - d1 * d6;
          if (vb <= 0.f && d2 >= 0.f && d6 <= 0.f) {
            const float w = d2 / (d2 - d6);
            q[0] = a[0] + w * ac[0]; q[1] = a[1] + w * ac[1]; q[2] = a[2] + w * ac[2];
          } else {
            const float va = d3 * d6 - d5 * d4;
            if (va <= 0.f && (d4 - d3) >= 0.f && (d5 - d6) >= 0.f) {
              const float w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
              q[0] = b[0] + w * (c[0] - b[0]); q[1] = b[1] + w * (c[1] - b[1]); q[2] = b[2] + w * (c[2] - b[2]);
            } else {
              const float den = 1.f / (va + vb + vc), v = vb * den, w = vc * den;
              q[0] = a[0] + ab[0] * v + ac[0] * w; q[1] = a[1] + ab[1] * v + ac[1] * w; q[2] = a[2] + ab[2] * v + ac[2] * w;
            }
          }
        }
      }
    }
  }
  const float dx = p[0] - q[0], dy = p[1] - q[1], dz = p[2] - q[2];
  return dx * dx + dy * dy + dz * dz;
}

__global__ void build_clearance_kernel(const float4 *__restrict__ tris, int n_tris, float ox, float oy, float oz, float h, int nx,
                                       int ny, int nz, float reach, unsigned *bits) {
  const int lane = threadIdx.x & 31;
  const int t = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (t >= n_tris) return;
  const float4 v0 = __ldg(tris + 3 * (size_t)t), v1 = __ldg(tris + 3 * (size_t)t + 1), v2 = __ldg(tris + 3 * (size_t)t + 2);
  const float a[3] = {v0.x, v0.y, v0.z}, b[3] = {v1.x, v1.y, v1.z}, c[3] = {v2.x, v2.y, v2.z};
  const float inv = 1.0f / h;
  int lo[3], hi[3];
  const float o[3] = {ox, oy, oz};
  const int n[3] = {nx, ny, nz};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float mn = fminf(a[k], fminf(b[k], c[k])) - reach, mx = fmaxf(a[k], fmaxf(b[k], c[k])) + reach;
    lo[k] = max(0, (int)floorf((mn - o[k]) * inv) - 1);
    hi[k] = min(n[k] - 1, (int)floorf((mx - o[k]) * inv) + 1);
  }
  const int sx = hi[0] - lo[0] + 1, sy = hi[1] - lo[1] + 1, sz = hi[2] - lo[2] + 1;
  if (sx <= 0 || sy <= 0 || sz <= 0) return;
  const long long total = (long long)sx * sy * sz;
  const float r2 = reach * reach;
  for (long long i = lane; i < total; i += 32) {
    const int ix = lo[0] + (int)(i % sx), iy = lo[1] + (int)((i / sx) % sy), iz = lo[2] + (int)(i / ((long long)sx * sy));
    const float p[3] = {ox + (ix + 0.5f) * h, oy + (iy + 0.5f) * h, oz + (iz + 0.5f) * h};
    if (!(point_tri_dist2(p, a, b, c) > r2)) {   // NaN (degenerate triangle) marks the cell: never optimistic
      const unsigned cell = ((unsigned)iz * (unsigned)ny + (unsigned)iy) * (unsigned)nx + (unsigned)ix;
      atomicOr(bits + (cell >> 5), 1u << (cell & 31));
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// synthetic pose stream: Philox4x32-10, bit-identical to oracle/sff_oracle.c::orc_gen_poses
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox10(unsigned long long seed, unsigned long long index, unsigned stream, unsigned *o) {
  unsigned c0 = (unsigned)index, c1 = (unsigned)(index >> 32), c2 = stream, c3 = 0;
  unsigned k0 = (unsigned)seed, k1 = (unsigned)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const unsigned n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}
__device__ __forceinline__ float u01(unsigned bits) { return __fmul_rn((float)(bits >> 8), 5.9604644775390625e-08f); }
__device__ __forceinline__ float acos_poly(float x) {
  const float ax = fabsf(x);
  float p = -0.0012624911f;
  p = __fadd_rn(__fmul_rn(p, ax), 0.0066700901f);
  p = __fadd_rn(__fmul_rn(p, ax), -0.0170881256f);
  p = __fadd_rn(__fmul_rn(p, ax), 0.0308918810f);
  p = __fadd_rn(__fmul_rn(p, ax), -0.0501743046f);
  p = __fadd_rn(__fmul_rn(p, ax), 0.0889789874f);
  p = __fadd_rn(__fmul_rn(p, ax), -0.2145988016f);
  p = __fadd_rn(__fmul_rn(p, ax), 1.5707963050f);
  const float r = __fmul_rn(__fsqrt_rn(__fsub_rn(1.0f, ax)), p);
  return x < 0.0f ? __fsub_rn(3.14159274f, r) : r;
}

__global__ void gen_poses_kernel(unsigned long long seed, unsigned long long first, long long n, float r0, float r1,
                                 float r2, float r3, float r4, float r5, float *out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned a[4], b[4];
  philox10(seed, first + (unsigned long long)i, 0, a);
  philox10(seed, first + (unsigned long long)i, 1, b);
  float *o = out + 6 * i;
  o[0] = __fadd_rn(r0, __fmul_rn(u01(a[0]), __fsub_rn(r1, r0)));
  o[1] = __fadd_rn(r2, __fmul_rn(u01(a[1]), __fsub_rn(r3, r2)));
  o[2] = __fadd_rn(r4, __fmul_rn(u01(a[2]), __fsub_rn(r5, r4)));
  o[3] = __fadd_rn(-3.14159274f, __fmul_rn(u01(a[3]), 6.28318548f));
  float phi = __fadd_rn(acos_poly(__fsub_rn(1.0f, __fmul_rn(2.0f, u01(b[0])))), 1.57079637f);
  if (u01(b[1]) < 0.5f) phi = __fsub_rn(phi, 3.14159274f);
  o[4] = phi;
  o[5] = __fadd_rn(-3.14159274f, __fmul_rn(u01(b[2]), 6.28318548f));
}


}  // namespace

// cudaFuncSetAttribute + occupancy are queried once per kernel instantiation (and per shared-memory size)
struct KernelCfg {
  size_t smem = ~size_t(0);
  int per_sm = 1;
  cudaError_t err = cudaSuccess;
};
template <auto Kernel>
static KernelCfg kernel_cfg(size_t smem) {
  static KernelCfg c;
  static std::mutex mu;   // environments of different host threads share the per-kernel cache
  std::lock_guard<std::mutex> lock(mu);
  if (c.smem != smem) {
    c.smem = smem;
    c.err = cudaFuncSetAttribute(Kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = 0;
    if (c.err == cudaSuccess &&
        (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, Kernel, kThreads, smem) != cudaSuccess || per_sm < 1))
      per_sm = 1;
    c.per_sm = per_sm < 1 ? 1 : per_sm;
  }
  return c;
}

size_t collide_smem_bytes(int n_robot, int n_stage_max) {
  return (size_t)n_robot * sizeof(RobotTri) + (size_t)2 * kTopSlots * sizeof(float4) +
         (size_t)n_stage_max * kWide * 2 * sizeof(float4) + (size_t)kWarpsPerBlock * sizeof(WarpScratch);
}


template <int FMT>
static cudaError_t launch_poses_fmt(const EnvDev &env, const void *d_poses, int64_t n, const OutSet &d_verdict, const GatherSync &gs,
                                    cudaStream_t stream, const LaunchCfg &cfg, bool count, int chunk, int *grid_out) {
  const size_t smem = collide_smem_bytes(env.n_robot, env.n_stage_max);
  const long long chunks = (n + chunk - 1) / chunk;
  const long long want = (chunks + kWarpsPerBlock - 1) / kWarpsPerBlock;
  cudaError_t e;
  if (count) {
    const KernelCfg kc = kernel_cfg<collide_poses_kernel<FMT, true>>(smem);
    if ((e = kc.err) != cudaSuccess) return e;
    int grid = cfg.sm_count * kc.per_sm;
    if (want < grid) grid = (int)want;
    *grid_out = grid;
    collide_poses_kernel<FMT, true><<<grid, kThreads, smem, stream>>>(env, d_poses, (long long)n, d_verdict, gs, chunk);
  } else {
    const KernelCfg kc = kernel_cfg<collide_poses_kernel<FMT, false>>(smem);
    if ((e = kc.err) != cudaSuccess) return e;
    int grid = cfg.sm_count * kc.per_sm;
    if (want < grid) grid = (int)want;
    *grid_out = grid;
    collide_poses_kernel<FMT, false><<<grid, kThreads, smem, stream>>>(env, d_poses, (long long)n, d_verdict, gs, chunk);
  }
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// cross-GPU completion barrier for the peer-store gather (one CTA of 32 threads, thread r talks to rank r)
// ---------------------------------------------------------------------------------------------------------
// signal == true : publish `epoch` to every rank (no wait)      -- what an empty shard owes its peers
// signal == false: wait until every rank has published >= epoch -- what a consumer of the gathered results enqueues
__global__ void peer_barrier_kernel(FlagSet f, unsigned epoch, bool signal, int *status) {
  if (signal) {
    const int r = threadIdx.x;
    if (r >= f.n) return;
    __threadfence_system();   // everything this stream did before is ordered before the signal
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f.p[r] + f.me), "r"(epoch) : "memory");
  } else if (threadIdx.x == 0) {
    wait_flags(f, epoch, status);
  }
}

cudaError_t launch_peer_barrier(const FlagSet &flags, unsigned epoch, bool signal, int *d_status, cudaStream_t stream) {
  peer_barrier_kernel<<<1, 32, 0, stream>>>(flags, epoch, signal, d_status);
  return cudaGetLastError();
}

// every warp of the grid leaves its loop through exactly one failing fetch, so a launch consumes
// (units + grid * warps) values of the shared counter; the host advances the base instead of resetting the counter
cudaError_t launch_collide_poses(const EnvDev &env_in, const void *d_poses, int pose_fmt, int64_t n, uint8_t *d_verdict,
                                 cudaStream_t stream, const LaunchCfg &cfg, bool count, unsigned *work_base_io) {
  OutSet one;
  one.n = 1;
  one.p[0] = d_verdict;
  GatherSync none{};
  return launch_collide_poses_gather(env_in, d_poses, pose_fmt, n, one, none, stream, cfg, count, work_base_io);
}

cudaError_t launch_collide_poses_gather(const EnvDev &env_in, const void *d_poses, int pose_fmt, int64_t n, const OutSet &d_verdict,
                                        const GatherSync &gs, cudaStream_t stream, const LaunchCfg &cfg, bool count,
                                        unsigned *work_base_io) {
  if (n <= 0) {
    // an empty shard still owes its peers the completion signal (after waiting like a real launch would)
    if (gs.flags.n > 0) {
      cudaError_t e = cudaSuccess;
      if (gs.wait_epoch != 0) e = launch_peer_barrier(gs.flags, gs.wait_epoch, false, env_in.status, stream);
      if (e == cudaSuccess) e = launch_peer_barrier(gs.flags, gs.signal_epoch, true, nullptr, stream);
      return e;
    }
    return cudaSuccess;
  }
  EnvDev env = env_in;
  env.work_base = *work_base_io;
  int grid = 0;
  // poses per work unit: enough units for ~2 per resident warp (at 2 CTAs/SM), capped at a full warp
  int chunk = 32;
  const long long warps = (long long)cfg.sm_count * 2 * kWarpsPerBlock * 2;
  while (chunk > 1 && (n + chunk - 1) / chunk < warps) chunk >>= 1;
  cudaError_t e = cudaErrorInvalidValue;
  switch (pose_fmt) {
    case 0: e = launch_poses_fmt<kFmtEulerF32>(env, d_poses, n, d_verdict, gs, stream, cfg, count, chunk, &grid); break;
    case 1: e = launch_poses_fmt<kFmtEulerF64>(env, d_poses, n, d_verdict, gs, stream, cfg, count, chunk, &grid); break;
    case 2: e = launch_poses_fmt<kFmtMatrixF64>(env, d_poses, n, d_verdict, gs, stream, cfg, count, chunk, &grid); break;
  }
  if (e == cudaSuccess) *work_base_io += (unsigned)((n + chunk - 1) / chunk) + (unsigned)grid * kWarpsPerBlock;
  return e;
}

cudaError_t launch_check_edges(const EnvDev &env_in, const double *d_starts, const double *d_ends, int64_t m,
                               double sample_dist, int rot_mode, uint8_t *d_free, int32_t *d_first_hit, cudaStream_t stream,
                               const LaunchCfg &cfg, bool count, unsigned *work_base_io, int *d_fh_scratch) {
  if (m <= 0) return cudaSuccess;
  EnvDev env = env_in;
  env.work_base = *work_base_io;
  const size_t smem = collide_smem_bytes(env.n_robot, env.n_stage_max);
  // warps per edge: enough units for ~2 per resident warp, at most 32 (needs the scratch array)
  int split = 1;
  const long long warps = (long long)cfg.sm_count * kWarpsPerBlock * SFFG_EDGE_UNITS_PER_WARP;
  if (d_fh_scratch)
    while (split < 32 && m * split < warps) split <<= 1;
  const long long units = m * split;
  const long long want = (units + kWarpsPerBlock - 1) / kWarpsPerBlock;
  cudaError_t e;
  if (split > 1 && (e = cudaMemsetAsync(d_fh_scratch, 0x7f, (size_t)m * sizeof(int), stream)) != cudaSuccess) return e;
  int grid;
  if (count) {
    const KernelCfg kc = kernel_cfg<check_edges_kernel<true>>(smem);
    if ((e = kc.err) != cudaSuccess) return e;
    grid = cfg.sm_count * kc.per_sm;
    if (want < grid) grid = (int)want;
    check_edges_kernel<true><<<grid, kThreads, smem, stream>>>(env, d_starts, d_ends, (long long)m, sample_dist, rot_mode, d_free, d_first_hit, split, d_fh_scratch);
  } else {
    const KernelCfg kc = kernel_cfg<check_edges_kernel<false>>(smem);
    if ((e = kc.err) != cudaSuccess) return e;
    grid = cfg.sm_count * kc.per_sm;
    if (want < grid) grid = (int)want;
    check_edges_kernel<false><<<grid, kThreads, smem, stream>>>(env, d_starts, d_ends, (long long)m, sample_dist, rot_mode, d_free, d_first_hit, split, d_fh_scratch);
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  *work_base_io += (unsigned)units + (unsigned)grid * kWarpsPerBlock;
  if (split > 1) {
    finalize_edges_kernel<<<(unsigned)((m + 255) / 256), 256, 0, stream>>>(d_fh_scratch, (long long)m, d_free, d_first_hit);
    e = cudaGetLastError();
  }
  return e;
}

cudaError_t launch_build_clearance(const float4 *d_tris32, int n_tris, const float origin[3], float h, const int n[3], float reach,
                                   unsigned *d_bits, cudaStream_t stream) {
  if (n_tris <= 0) return cudaSuccess;
  const int threads = 256;
  const long long blocks = ((long long)n_tris * 32 + threads - 1) / threads;
  build_clearance_kernel<<<(unsigned)blocks, threads, 0, stream>>>(d_tris32, n_tris, origin[0], origin[1], origin[2], h, n[0], n[1],
                                                                    n[2], reach, d_bits);
  return cudaGetLastError();
}

cudaError_t launch_gen_poses(uint64_t seed, uint64_t first, int64_t n, const float range[6], float *d_out,
                             cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  const int threads = 256;
  const long long blocks = (n + threads - 1) / threads;
  gen_poses_kernel<<<(unsigned)blocks, threads, 0, stream>>>(seed, first, (long long)n, range[0], range[1], range[2],
                                                             range[3], range[4], range[5], d_out);
  return cudaGetLastError();
}

}  // namespace sffg
