// Host-side refinement that follows the GPU hot path (BASELINE.json north_star: "The Hungarian /
// munkres permutation step and the Kabsch refinement stay on the host"), as native C++ on an OpenMP
// thread pool over independent pairs:
//   * linear assignment: dense shortest-augmenting-path Jonker-Volgenant (the role of JOVOSAP,
//     reference fastoverlap/f90/alignutils.f90:989-1259, and of munkres / pele in utils.py:34-60);
//     same algorithm and tie-breaking as scipy.optimize.linear_sum_assignment, which the Python
//     classes of this package use, so both host paths return the same permutations;
//   * periodic: permutation <-> mean-displacement iteration (periodicAlignment.py:27-80,
//     FINDDISPLACEMENT alignutils.f90:381-419);
//   * clusters: rotate by the Euler angles of the grid maximum (utils.py:447-460), permute, Kearsley
//     quaternion fit (utils.py:169-253, FINDROTATION alignutils.f90:304-379), best orientation by
//     distance (the Fortran rule, fastclusters.f90:243-254).
// No CUDA in this file; it is compiled into the same shared library.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <limits>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "../../include/fastoverlap_b200.h"

namespace {

// fo_host_refine_counters: [0] LAP solves, [1] screened assignments, [2] skipped repeat solves
long long g_counters[3] = {0, 0, 0};
inline void count(int k) { __atomic_fetch_add(&g_counters[k], 1LL, __ATOMIC_RELAXED); }

// ---------------------------------------------------------------------------------------------------------
// Inner loops of the assignment, written with GCC vector extensions (one vector of the widest native type per
// statement; a 64-byte vector type on a narrower ISA is scalarised, so every kernel is stamped out per ISA and
// picked once at load time).  The auto-vectorised forms of these loops kept their running minima on the stack
// and ran at 4-5 cycles per matrix entry; the search for the minimising column was a scalar pass with a branch
// per column and took most of a hard assignment (profiles/r02_summary.md, "Host pool").
//
// lap_pad(n): row pitch of the cost matrix and length of the padded column arrays (a multiple of the widest
// vector).  Padding columns repeat the last real column; nothing reads their results.
constexpr int LAPW = 8;
inline int lap_pad(int n) { return (n + LAPW - 1) / LAPW * LAPW; }

// Horizontal (value, index) minimum of a vector pair: smallest value, on ties the smallest index, as log2(W)
// shuffle / compare / blend steps (the scalar form compiles into a data-dependent branch per lane, and the lane
// of the minimum is as good as random).  V / I are overwritten; afterwards every lane holds the result.
#define FO_HSTEP(V, I, ...)                                          \
  {                                                                  \
    const vd sv_ = __builtin_shufflevector(V, V, __VA_ARGS__);       \
    const vd si_ = __builtin_shufflevector(I, I, __VA_ARGS__);       \
    const vl t_ = (sv_ < V) | ((sv_ == V) & (si_ < I));              \
    V = t_ ? sv_ : V;                                                \
    I = t_ ? si_ : I;                                                \
  }
#define FO_HMIN8(V, I) FO_HSTEP(V, I, 4, 5, 6, 7, 0, 1, 2, 3) FO_HSTEP(V, I, 2, 3, 0, 1, 6, 7, 4, 5) FO_HSTEP(V, I, 1, 0, 3, 2, 5, 4, 7, 6)
#define FO_HMIN4(V, I) FO_HSTEP(V, I, 2, 3, 0, 1) FO_HSTEP(V, I, 1, 0, 3, 2)
#define FO_HMIN2(V, I) FO_HSTEP(V, I, 1, 0)

// FO_LAP_KERNELS(S, W, HMIN, ATTR) defines, for vectors of W doubles:
//  scan_S   one row scan of the shortest-augmenting-path search (Lap::solve): relaxes every column through row
//           i, returns the smallest tentative distance and picks its column -- on ties an unassigned column
//           first, then the lowest index (scipy's rule).  cd: tentative distances (+inf for closed columns),
//           cl: 0 / +inf for open / closed columns, asg: 0 / +inf for unassigned / assigned columns, pd:
//           predecessor row (as a double, one lane type).  The arithmetic of r is scipy's, term by term.  All
//           column arrays have nl = lap_pad(n) entries; v = -inf in the padding keeps those columns at +inf.
//  cost_S   squared (minimum-image, box != null) distances between the rows xs and the columns ys (structure
//           of arrays, ys padded to lap_pad(n)), with the running minimum of every column, its row and the
//           second smallest entry (the column reduction of the LAP and the gap test of best_perm) kept in
//           registers over the sweep down the rows.  rint(t) = (t + 1.5 * 2^52) - 1.5 * 2^52 in round-to-nearest.
#define FO_LAP_KERNELS(S, W, HMIN, ATTR)                                                                       \
  ATTR double scan_##S(int nl, const double* ci, const double* v, const double* cl, const double* asg,        \
                       double* cd, double* pd, double minVal, double ui, double di, int* pick) {              \
    typedef double vd __attribute__((vector_size(8 * W)));                                                    \
    typedef long long vl __attribute__((vector_size(8 * W)));                                                 \
    const double inf = std::numeric_limits<double>::infinity();                                               \
    vd vlo = (vd){} + inf, vlou = vlo, vli = (vd){}, vliu = (vd){}, jv;                                       \
    {                                                                                                         \
      double t_[W];                                                                                           \
      for (int l = 0; l < W; ++l) t_[l] = (double)l;                                                          \
      memcpy(&jv, t_, sizeof(vd));                                                                            \
    }                                                                                                         \
    for (int j = 0; j < nl; j += W) {                                                                         \
      vd c_, v_, l_, d_, p_, a_;                                                                              \
      memcpy(&c_, ci + j, sizeof(vd));                                                                        \
      memcpy(&v_, v + j, sizeof(vd));                                                                         \
      memcpy(&l_, cl + j, sizeof(vd));                                                                        \
      memcpy(&d_, cd + j, sizeof(vd));                                                                        \
      memcpy(&p_, pd + j, sizeof(vd));                                                                        \
      memcpy(&a_, asg + j, sizeof(vd));                                                                       \
      const vd r = minVal + c_ - ui - v_ + l_;                                                                \
      const vl lt = r < d_;                                                                                   \
      const vd c = lt ? r : d_;                                                                               \
      p_ = lt ? (vd){} + di : p_;                                                                             \
      memcpy(cd + j, &c, sizeof(vd));                                                                         \
      memcpy(pd + j, &p_, sizeof(vd));                                                                        \
      const vl m = c < vlo;                                                                                   \
      vli = m ? jv : vli;                                                                                     \
      vlo = m ? c : vlo;                                                                                      \
      const vd cu = c + a_;                                                                                   \
      const vl mu = cu < vlou;                                                                                \
      vliu = mu ? jv : vliu;                                                                                  \
      vlou = mu ? cu : vlou;                                                                                  \
      jv += (double)W;                                                                                        \
    }                                                                                                         \
    HMIN(vlo, vli)                                                                                            \
    HMIN(vlou, vliu)                                                                                          \
    const double lo = vlo[0], lou = vlou[0], li = vli[0], liu = vliu[0];                                      \
    *pick = lo == inf ? -1 : (lou == lo ? (int)liu : (int)li);                                                \
    return lo;                                                                                                \
  }                                                                                                           \
  ATTR void cost_##S(int n, int ld, const double* xs, const double* ys, const double* box, double* cost,      \
                     double* vmin, int* imin, double* vmin2) {                                                \
    typedef double vd __attribute__((vector_size(8 * W)));                                                    \
    typedef long long vl __attribute__((vector_size(8 * W)));                                                 \
    const double inf = std::numeric_limits<double>::infinity(), M = 6755399441055744.0;                       \
    const double b0 = box ? box[0] : 1.0, b1 = box ? box[1] : 1.0, b2 = box ? box[2] : 1.0;                   \
    const double i0 = 1.0 / b0, i1 = 1.0 / b1, i2 = 1.0 / b2;                                                 \
    for (int jb = 0; jb < ld; jb += W) {                                                                      \
      vd y0, y1, y2;                                                                                          \
      memcpy(&y0, ys + jb, sizeof(vd));                                                                       \
      memcpy(&y1, ys + ld + jb, sizeof(vd));                                                                  \
      memcpy(&y2, ys + 2 * ld + jb, sizeof(vd));                                                              \
      vd vm = (vd){} + inf, v2 = vm, im = (vd){};                                                             \
      if (box) {                                                                                              \
        for (int i = 0; i < n; ++i) {                                                                         \
          vd dx = xs[i] - y0, dy = xs[n + i] - y1, dz = xs[2 * n + i] - y2;                                   \
          dx -= ((dx * i0 + M) - M) * b0;                                                                     \
          dy -= ((dy * i1 + M) - M) * b1;                                                                     \
          dz -= ((dz * i2 + M) - M) * b2;                                                                     \
          const vd d = dx * dx + dy * dy + dz * dz;                                                           \
          memcpy(cost + (size_t)i * ld + jb, &d, sizeof(vd));                                                 \
          const vl lt = d < vm;                                                                               \
          const vd hi = lt ? vm : d; /* the larger of the entry and the running minimum */                    \
          v2 = hi < v2 ? hi : v2;                                                                             \
          vm = lt ? d : vm;                                                                                   \
          im = lt ? (vd){} + (double)i : im;                                                                  \
        }                                                                                                     \
      } else {                                                                                                \
        for (int i = 0; i < n; ++i) {                                                                         \
          const vd dx = xs[i] - y0, dy = xs[n + i] - y1, dz = xs[2 * n + i] - y2;                             \
          const vd d = dx * dx + dy * dy + dz * dz;                                                           \
          memcpy(cost + (size_t)i * ld + jb, &d, sizeof(vd));                                                 \
          const vl lt = d < vm;                                                                               \
          const vd hi = lt ? vm : d;                                                                          \
          v2 = hi < v2 ? hi : v2;                                                                             \
          vm = lt ? d : vm;                                                                                   \
          im = lt ? (vd){} + (double)i : im;                                                                  \
        }                                                                                                     \
      }                                                                                                       \
      double tv_[W], t2_[W], ti_[W];                                                                          \
      memcpy(tv_, &vm, sizeof(tv_));                                                                          \
      memcpy(t2_, &v2, sizeof(t2_));                                                                          \
      memcpy(ti_, &im, sizeof(ti_));                                                                          \
      for (int l = 0; l < W && jb + l < n; ++l) {                                                             \
        vmin[jb + l] = tv_[l];                                                                                \
        vmin2[jb + l] = t2_[l];                                                                               \
        imin[jb + l] = (int)ti_[l];                                                                           \
      }                                                                                                       \
    }                                                                                                         \
  }                                                                                                           \
  ATTR void sqrt_##S(int n, double* c) {                                                                      \
    for (int j = 0; j < n; ++j) c[j] = __builtin_sqrt(c[j]);                                                  \
  }

struct LapKernels {
  double (*scan)(int, const double*, const double*, const double*, const double*, double*, double*, double, double,
                 double, int*);
  void (*cost)(int, int, const double*, const double*, const double*, double*, double*, int*, double*);
  void (*sqrt_row)(int, double*);
};
// isa: 1 = 2-wide vectors (SSE2 / generic), 2 = AVX2, 3 = AVX-512; 0 = the widest the CPU has.  -> false: not available
#if defined(__GNUC__) && defined(__x86_64__) && !defined(__CUDACC__)
FO_LAP_KERNELS(avx512, 8, FO_HMIN8, __attribute__((target("avx512f"))))
FO_LAP_KERNELS(avx2, 4, FO_HMIN4, __attribute__((target("avx2"))))
FO_LAP_KERNELS(base, 2, FO_HMIN2, )
bool pick_lap_kernels(int isa, LapKernels* k, int* picked) {
  __builtin_cpu_init();
  const bool a512 = __builtin_cpu_supports("avx512f"), a2 = __builtin_cpu_supports("avx2");
  if (isa == 0) isa = a512 ? 3 : (a2 ? 2 : 1);
  if (isa == 3 && a512) *k = {scan_avx512, cost_avx512, sqrt_avx512};
  else if (isa == 2 && a2) *k = {scan_avx2, cost_avx2, sqrt_avx2};
  else if (isa == 1) *k = {scan_base, cost_base, sqrt_base};
  else return false;
  *picked = isa;
  return true;
}
#else
FO_LAP_KERNELS(base, 2, FO_HMIN2, )
bool pick_lap_kernels(int isa, LapKernels* k, int* picked) {
  if (isa > 1) return false;
  *k = {scan_base, cost_base, sqrt_base};
  *picked = 1;
  return true;
}
#endif
LapKernels LK;
int g_lap_isa = 0;
const bool g_lap_init = pick_lap_kernels(0, &LK, &g_lap_isa);

// Dense n x n linear assignment (minimise), cost row-major with pitch ld.  col4row[i] = column assigned to row i.
struct Lap {
  std::vector<double> u, v, shortest, cand, closed, pathd, asg;
  std::vector<int> path, row4col, colmin, srows, scols;
  std::vector<char> rowdone;
  // colmin_in / vmin_in (optional): per-column minimum and its row, when the caller already has them.
  // lazy_sqrt: cost (and vmin_in) hold SQUARED costs; the LAP runs on their square roots, taken row by row
  // only for the rows an augmentation actually scans.  The column reduction needs no roots but those of
  // the column minima (sqrt is monotone), and after a good alignment it already assigns almost every row,
  // so a 204 x 204 periodic cost matrix costs ~200 square roots instead of 41616 (the vector sqrt was
  // 2/3 of the whole host refinement of a BLJ256 pair).
  // ld = lap_pad(n): the column arrays are padded to whole vectors (v = -inf there: a padding column stays
  // at +inf in every scan), the matrix rows are read up to ld.
  void solve(int n, int ld, double* cost, int* col4row, const int* colmin_in = nullptr,
             const double* vmin_in = nullptr, bool lazy_sqrt = false) {
    if (lazy_sqrt) rowdone.assign(n, 0);
    auto row_of = [&](int i) -> const double* {
      double* ci = cost + (size_t)i * ld;
      if (lazy_sqrt && !rowdone[i]) {
        LK.sqrt_row(n, ci);
        rowdone[i] = 1;
      }
      return ci;
    };
    u.assign(n, 0.0);
    v.assign(ld, -std::numeric_limits<double>::infinity());
    shortest.resize(n);
    path.resize(n);
    row4col.assign(n, -1);
    for (int i = 0; i < n; ++i) col4row[i] = -1;
    const double inf = std::numeric_limits<double>::infinity();
    // Column reduction (the initialisation phase of Jonker-Volgenant, JOVOSAP alignutils.f90:1041-1075):
    // v[j] = min_i c[i][j], and column j is given to its minimising row when that row is still free.
    // With u = 0 this is dual feasible and tight on the assigned pairs, so the augmentation below only
    // has to run for the rows left free -- after a good alignment that is almost none of them.
    {
      const int* cm = colmin_in;
      if (cm) {
        for (int j = 0; j < n; ++j) v[j] = lazy_sqrt ? __builtin_sqrt(vmin_in[j]) : vmin_in[j];
      } else {
        colmin.assign(n, 0);
        for (int j = 0; j < n; ++j) v[j] = row_of(0)[j];  // (all of v[0 .. n) is set on both branches)
        for (int i = 1; i < n; ++i) {
          const double* ci = row_of(i);
          for (int j = 0; j < n; ++j)
            if (ci[j] < v[j]) {
              v[j] = ci[j];
              colmin[j] = i;
            }
        }
        cm = colmin.data();
      }
      for (int j = n - 1; j >= 0; --j) {
        const int i = cm[j];
        if (v[j] == v[j] && col4row[i] == -1) {
          col4row[i] = j;
          row4col[j] = i;
        }
      }
    }
    // (The augmenting row reduction of Jonker-Volgenant, JOVOSAP alignutils.f90:1077-1135, was used here in
    // round 1 for partially aligned pairs; with the vector scan below a search step costs what a reduction step
    // costs and the reduction needs more of them -- LJ38 inverted orientation 18.4 -> 11.3 us, BLJ256 at a
    // jitter of 0.3: 328 -> 234 us per pair without it -- so the free rows go straight to the search.)
    // Shortest augmenting path per free row.  The scan of a row runs over ALL columns without branches or
    // index indirection (LK.scan): cand[j] is the tentative distance of an open column and +inf once the
    // column is closed (then shortest[j] keeps its final distance for the dual update), closed[j] = +inf keeps
    // a closed column from being reopened, asg[j] = +inf marks the assigned columns for the tie rule.  The rows
    // and columns an augmentation touches are kept as lists: the dual update costs the length of the search
    // tree, not n.
    cand.resize(ld);
    closed.resize(ld);
    pathd.resize(ld);
    asg.assign(ld, inf);
    for (int j = 0; j < n; ++j) asg[j] = row4col[j] == -1 ? 0.0 : inf;
    for (int cur = 0; cur < n; ++cur) {
      if (col4row[cur] != -1) continue;
      srows.clear();
      scols.clear();
      double* cd = cand.data();
      double* cl = closed.data();
      double* pd = pathd.data();
      const double* vv = v.data();
      for (int j = 0; j < ld; ++j) {
        cd[j] = inf;
        cl[j] = 0.0;
      }
      int sink = -1, i = cur;
      double minVal = 0.0;
      while (sink == -1) {
        srows.push_back(i);
        const double* ci = row_of(i);
        int j;
        const double lowest = LK.scan(ld, ci, vv, cl, asg.data(), cd, pd, minVal, u[i], (double)i, &j);
        if (lowest == inf || j < 0) return;  // infeasible (NaN costs): leave -1s
        minVal = lowest;
        shortest[j] = lowest;
        path[j] = (int)pd[j];
        cd[j] = inf;
        cl[j] = inf;
        scols.push_back(j);
        if (row4col[j] == -1)
          sink = j;
        else
          i = row4col[j];
      }
      u[cur] += minVal;
      for (size_t k = 1; k < srows.size(); ++k) {
        const int r = srows[k];
        u[r] += minVal - shortest[col4row[r]];
      }
      for (const int j : scols) v[j] -= minVal - shortest[j];
      asg[sink] = inf;
      int j = sink;
      while (true) {
        const int r = path[j];
        row4col[j] = r;
        std::swap(col4row[r], j);
        if (r == cur) break;
      }
    }
  }
};

// rint == nearbyint in the default rounding mode; it inlines to one roundsd (-msse4.1).  d * (1 / box)
// instead of d / box (a division costs ~4 cycles of throughput and the pair loop takes ~3800 minimum images)
// can pick the other image only within rounding of exactly half a box, where both are equally near.
inline double min_image(double d, double box, double ibox) { return d - __builtin_rint(d * ibox) * box; }

// Canonical summation order shared with the device path (fo_assign.cu: canon_sum): 32 strided partial sums,
// each accumulated in index order, combined pairwise by an xor butterfly.  The displacement update and the final
// distance use it so that pairs settled by the device screening and pairs settled here agree bit for bit
// (this translation unit is compiled without FMA contraction targets for these loops: plain mul / add).
inline double canon_sum(const double* t, int n) {
  double p[32], q[32];
  for (int l = 0; l < 32; ++l) {
    double s = 0.0;
    for (int i = l; i < n; i += 32) s += t[i];
    p[l] = s;
  }
  for (int off = 16; off; off >>= 1) {
    for (int l = 0; l < 32; ++l) q[l] = p[l] + p[l ^ off];
    for (int l = 0; l < 32; ++l) p[l] = q[l];
  }
  return p[0];
}

// permutation groups as the C ABI passes them; validated like fo_set_perm does for the device path
inline bool groups_valid(const int32_t* goff, int64_t ngroups, const int32_t* gidx, int64_t natoms) {
  if (ngroups < 1 || !goff || !gidx || natoms < 1 || goff[0] != 0) return false;
  for (int64_t g = 0; g < ngroups; ++g)
    if (goff[g + 1] < goff[g]) return false;
  if (goff[ngroups] > natoms) return false;
  for (int64_t i = 0; i < goff[ngroups]; ++i)
    if (gidx[i] < 0 || gidx[i] >= natoms) return false;
  return true;
}

struct Groups {
  const int32_t* goff;
  int64_t ngroups;
  const int32_t* gidx;
};

// Cost matrices of one permutation group from structure-of-arrays coordinates; the inner loops are
// branch-free so that gcc vectorises them (function multiversioning picks the AVX2 clone at run time).
// d * (1/box) instead of d / box can move the rounding only at exact half-box separations, where both
// images give the same distance.

// Single-precision screening pass of the periodic assignment: smallest and second smallest min-image
// distance (squared) of every column and the row of the smallest, 16 columns at a time, nothing stored.
// After a good alignment every atom's partner is far closer than anything else, the column minima form a
// permutation, and that permutation is the unique optimum of the assignment -- proven by best_perm from
// these three arrays with the rounding error of this pass as a tolerance, so that neither the double-
// precision matrix nor the LAP is needed.  Coordinates must be wrapped into [-box/2, box/2] (error analysis
// in best_perm).
// Written with GCC vector extensions, one vector of the widest native type per statement (the scalar form of
// this update was compiled into mask tests and branches with the running minima spilled to the stack; a
// 64-byte vector type on a narrower ISA is scalarised), stamped out per ISA and picked once at run time.
// rint(t) = (t + 1.5 * 2^23) - 1.5 * 2^23 for |t| < 2^22 in round-to-nearest.  ys is padded to a multiple
// of CBF columns per component (pitch npad), the padding repeating a real column.  Every block of CBF columns
// visits two runs of rows only (runs: 4 ints per block), see screen_periodic.
constexpr int CBF = 16;

#define FO_COLMIN_F32(NAME, W, ATTR)                                                                          \
  ATTR void NAME(int n, int npad, const float* xs, const float* ys, const float* box, const int* runs,         \
                 float* vmin, int* imin, float* vmin2) {                                                       \
    typedef float vf __attribute__((vector_size(4 * W)));                                                      \
    typedef int vi __attribute__((vector_size(4 * W)));                                                        \
    const float b0 = box[0], b1 = box[1], b2 = box[2];                                                         \
    const float i0 = 1.0f / b0, i1 = 1.0f / b1, i2 = 1.0f / b2;                                                \
    const float inf = std::numeric_limits<float>::infinity(), M = 12582912.0f;                                 \
    for (int jb = 0; jb < npad; jb += W) {                                                                     \
      vf y0, y1, y2;                                                                                           \
      memcpy(&y0, ys + jb, sizeof(vf));                                                                        \
      memcpy(&y1, ys + npad + jb, sizeof(vf));                                                                 \
      memcpy(&y2, ys + 2 * npad + jb, sizeof(vf));                                                             \
      vf vm = y0 * 0.0f + inf, v2 = vm;                                                                        \
      vi im = {0};                                                                                             \
      const int* rn = runs + 4 * (jb / CBF); /* rows [rn[0], rn[1]) and [rn[2], rn[3]) */                      \
      for (int i = rn[0]; i < rn[3]; ++i) {                                                                    \
        if (i == rn[1]) {                                                                                      \
          i = rn[2];                                                                                           \
          if (i >= rn[3]) break;                                                                               \
        }                                                                                                      \
        vf dx = xs[i] - y0, dy = xs[n + i] - y1, dz = xs[2 * n + i] - y2;                                      \
        dx -= ((dx * i0 + M) - M) * b0;                                                                        \
        dy -= ((dy * i1 + M) - M) * b1;                                                                        \
        dz -= ((dz * i2 + M) - M) * b2;                                                                        \
        const vf d = dx * dx + dy * dy + dz * dz;                                                              \
        const vi lt = d < vm;                                                                                  \
        const vf hi = lt ? vm : d; /* the larger of the entry and the running minimum */                       \
        v2 = hi < v2 ? hi : v2;                                                                                \
        vm = lt ? d : vm;                                                                                      \
        im = lt ? (vi){0} + i : im;                                                                            \
      }                                                                                                        \
      memcpy(vmin + jb, &vm, sizeof(vf));                                                                      \
      memcpy(vmin2 + jb, &v2, sizeof(vf));                                                                     \
      memcpy(imin + jb, &im, sizeof(vi));                                                                      \
    }                                                                                                          \
  }

#if defined(__GNUC__) && defined(__x86_64__) && !defined(__CUDACC__)
FO_COLMIN_F32(colmin_f32_avx512, 16, __attribute__((target("avx512f"))))
FO_COLMIN_F32(colmin_f32_avx2, 8, __attribute__((target("avx2,fma"))))
FO_COLMIN_F32(colmin_f32_base, 4, )
typedef void (*colmin_f32_fn)(int, int, const float*, const float*, const float*, const int*, float*, int*,
                              float*);
colmin_f32_fn pick_colmin_f32() {
  __builtin_cpu_init();
  if (__builtin_cpu_supports("avx512f")) return colmin_f32_avx512;
  if (__builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma")) return colmin_f32_avx2;
  return colmin_f32_base;
}
const colmin_f32_fn colmin_periodic_f32 = pick_colmin_f32();
#else
FO_COLMIN_F32(colmin_periodic_f32, 4, )
#endif

// Screening of one periodic group in single precision (colmin_periodic_f32).  Returns true, the assignment
// (c4r[row] = column) and the updated stability margin when the column minima provably form the unique
// optimal assignment; false when the double-precision matrix + LAP has to decide.
// Rounding: wrapped coordinates |x| <= b / 2 (b = largest box length) are cast with an error <= 3e-8 b; the
// difference, the image shift and the rounding of the box lengths add <= 7e-7 b per component in the worst
// case (a difference within rounding of half a box may take the other image, which moves it by twice its
// distance to the half box), so a distance is off by less than 2e-6 b.  tol = 1e-5 b is used: row i_j is
// certainly the strict minimum of column j when sqrt(second) - sqrt(first) > 2 tol in single precision, and
// the true gap is at least that difference - 2 tol.
bool screen_periodic(int n, const double* xs, const double* ys, int ldy, const double* box, int* imin,
                     int* c4r, double* margin) {
  static thread_local std::vector<float> f;
  static thread_local std::vector<int> iw;
  static thread_local std::vector<char> seen;
  const int npad = (n + CBF - 1) / CBF * CBF, nblk = npad / CBF;
  constexpr int SMAX = 64;
  if (f.size() < (size_t)3 * n + 5 * npad + 2 * n) {
    f.resize((size_t)3 * n + 5 * npad + 2 * n);
    iw.resize((size_t)npad + 4 * n + 4 * nblk + 2 * (SMAX + 1));
  }
  float* xf = f.data();
  float* yf = xf + 3 * n;
  float* v1 = yf + 3 * npad;
  float* v2 = v1 + npad;
  float* wx = v2 + npad;  // wrapped coordinate along the slab axis, rows / columns
  float* wy = wx + n;
  int* im = iw.data();
  int* xord = im + npad;  // sorted position -> row / column of the group
  int* yord = xord + n;
  int* sx = yord + n;     // slab of every row / column
  int* sy = sx + n;
  int* runs = sy + n;
  int* xstart = runs + 4 * nblk;  // first sorted row of every slab
  int* ystart = xstart + SMAX + 1;
  const float boxf[3] = {(float)box[0], (float)box[1], (float)box[2]};
  const double bmax = std::max(box[0], std::max(box[1], box[2]));
  const double tol = 1e-5 * bmax;
  // Slabs along the longest axis.  A column only needs the rows within c = half the mean spacing of the group:
  // a partner further away than that is not an unambiguous nearest neighbour anyway.  Rows and columns are
  // bucket-sorted by slab; the 16 columns of a block span slabs sa..sb, and rows outside sa-k..sb+k (cyclic)
  // are at least ceff = k w - tol away from all of them (w: slab width), so they are skipped and the second
  // minimum is capped at ceff: the gap test below stays a proof.
  const int ax = box[0] >= box[1] ? (box[0] >= box[2] ? 0 : 2) : (box[1] >= box[2] ? 1 : 2);
  const double c = 0.5 * cbrt(box[0] * box[1] * box[2] / n);
  int S = (int)(box[ax] / (c / 3.0));
  S = S > SMAX ? SMAX : S;
  int kk = 0;
  double ceff2 = std::numeric_limits<double>::infinity();
  if (S >= 8) {
    const double w = box[ax] / S;
    kk = (int)ceil((c + tol) / w);
    const double ceff = kk * w - tol;
    ceff2 = ceff * ceff;
  }
  const bool slabs = S >= 8 && 2 * kk + 2 < S;
  {
    const double b = box[ax], ib = 1.0 / b;
    for (int i = 0; i < n; ++i) {
      const double x = xs[ax * n + i], y = ys[ax * ldy + i];
      wx[i] = (float)(x - __builtin_rint(x * ib) * b);
      wy[i] = (float)(y - __builtin_rint(y * ib) * b);
    }
  }
  if (slabs) {
    const float fs = (float)S, ibf = 1.0f / boxf[ax];
    for (int s = 0; s <= S; ++s) xstart[s] = ystart[s] = 0;
    for (int i = 0; i < n; ++i) {
      int a = (int)((wx[i] * ibf + 0.5f) * fs), b = (int)((wy[i] * ibf + 0.5f) * fs);
      a = a < 0 ? 0 : (a >= S ? S - 1 : a);
      b = b < 0 ? 0 : (b >= S ? S - 1 : b);
      sx[i] = a;
      sy[i] = b;
      ++xstart[a + 1];
      ++ystart[b + 1];
    }
    for (int s = 0; s < S; ++s) {
      xstart[s + 1] += xstart[s];
      ystart[s + 1] += ystart[s];
    }
    // counting sort; xstart / ystart are advanced while placing and restored afterwards
    for (int i = 0; i < n; ++i) {
      xord[xstart[sx[i]]++] = i;
      yord[ystart[sy[i]]++] = i;
    }
    for (int s = S; s > 0; --s) xstart[s] = xstart[s - 1];
    xstart[0] = 0;
    for (int b = 0; b < nblk; ++b) {
      const int jl = std::min(b * CBF + CBF - 1, n - 1);
      const int sa = sy[yord[b * CBF]], sb = sy[yord[jl]];
      int* rn = runs + 4 * b;
      const int lo = sa - kk, hi = sb + kk;
      if (hi - lo + 1 >= S) {
        rn[0] = 0; rn[1] = n; rn[2] = n; rn[3] = n;
      } else if (lo < 0) {
        rn[0] = 0; rn[1] = xstart[hi + 1]; rn[2] = xstart[lo + S]; rn[3] = n;
      } else if (hi >= S) {
        rn[0] = 0; rn[1] = xstart[hi - S + 1]; rn[2] = xstart[lo]; rn[3] = n;
      } else {
        rn[0] = xstart[lo]; rn[1] = xstart[hi + 1]; rn[2] = n; rn[3] = n;
      }
    }
  } else {
    for (int i = 0; i < n; ++i) xord[i] = yord[i] = i;
    for (int b = 0; b < nblk; ++b) {
      int* rn = runs + 4 * b;
      rn[0] = 0; rn[1] = n; rn[2] = n; rn[3] = n;
    }
    ceff2 = std::numeric_limits<double>::infinity();
  }
  for (int k = 0; k < 3; ++k) {
    const double b = box[k], ib = 1.0 / b;
    for (int i = 0; i < n; ++i) {
      const double x = xs[k * n + xord[i]], y = ys[k * ldy + yord[i]];
      xf[k * n + i] = (float)(x - __builtin_rint(x * ib) * b);
      yf[k * npad + i] = (float)(y - __builtin_rint(y * ib) * b);
    }
    for (int i = n; i < npad; ++i) yf[k * npad + i] = yf[k * npad + n - 1];
  }
  colmin_periodic_f32(n, npad, xf, yf, boxf, runs, v1, im, v2);
  const double tol2 = 2.0 * tol;
  seen.assign(n, 0);
  double gap = std::numeric_limits<double>::infinity();
  for (int j = 0; j < n; ++j) {
    const double second = std::min((double)v2[j], ceff2);
    const double g = __builtin_sqrt(second) - __builtin_sqrt((double)v1[j]) - tol2;
    const int row = im[j];
    if (!(g > 0) || row < 0 || row >= n || seen[row]) return false;  // also catches NaN coordinates, empty runs
    seen[row] = 1;
    gap = std::min(gap, g);
    imin[yord[j]] = xord[row];
    c4r[xord[row]] = yord[j];
  }
  *margin = std::min(*margin, gap);
  return true;
}

// permutation of Y that best matches X group by group; periodic (box != null: cost = min-image
// distance, periodicAlignment.py:94-102) or free (squared distance, utils.py:48-56).
// Returns the stability margin of the periodic result: when in every group every column's smallest entry
// sits in a different row, that assignment is the unique optimum (it attains the lower bound sum_j min_i
// c_ij), and it stays the unique optimum for as long as no column minimum changes rows.  The min-image
// distance is a metric on the torus, so moving all of Y by delta changes every entry by at most |delta|:
// the permutation provably survives any |delta| < margin / 2, margin = min_j (second smallest - smallest
// entry of column j).  -1 when the column minima do not form a permutation (or for free costs).
// screen (optional, one flag per group, in / out): whether the single-precision screening is worth trying;
// cleared for a group once it fails, so that the later solves of a badly aligned pair go straight to the LAP.
double best_perm(const Groups& G, int natoms, const double* X, const double* Y, const double* box, Lap& lap,
                 std::vector<double>& cost, std::vector<int>& c4r, int* perm, char* screen = nullptr) {
  static thread_local std::vector<double> soa, vmin, vmin2;
  static thread_local std::vector<int> imin;
  static thread_local std::vector<char> seen;
  double margin = box ? std::numeric_limits<double>::infinity() : -1.0;
  for (int i = 0; i < natoms; ++i) perm[i] = i;
  for (int64_t g = 0; g < G.ngroups; ++g) {
    const int n = G.goff[g + 1] - G.goff[g];
    if (n == 0) continue;
    const int32_t* idx = G.gidx + G.goff[g];
    // grow only: shrinking for a small group and growing again would zero-fill the matrix every time
    const int ld = lap_pad(n);  // row pitch of the matrix = padded number of columns
    if (cost.size() < (size_t)n * ld) cost.resize((size_t)n * ld);
    if (c4r.size() < (size_t)n) c4r.resize(n);
    if (soa.size() < (size_t)3 * n + 3 * ld) soa.resize((size_t)3 * n + 3 * ld);
    double* xs = soa.data();
    double* ys = xs + 3 * n;  // pitch ld, the padding repeats the last column
    for (int i = 0; i < n; ++i)
      for (int k = 0; k < 3; ++k) {
        xs[k * n + i] = X[3 * idx[i] + k];
        ys[k * ld + i] = Y[3 * idx[i] + k];
      }
    for (int i = n; i < ld; ++i)
      for (int k = 0; k < 3; ++k) ys[k * ld + i] = ys[k * ld + n - 1];
    if (vmin.size() < (size_t)n) {
      vmin.resize(n);
      vmin2.resize(n);
      imin.resize(n);
    }
    if (box && screen && screen[g] && n >= 2 && n < (1 << 24)) {
      if (screen_periodic(n, xs, ys, ld, box, imin.data(), c4r.data(), &margin)) {
        for (int i = 0; i < n; ++i) perm[idx[i]] = idx[c4r[i]];
        count(1);
        continue;
      }
      screen[g] = 0;
    }
    if (box) count(0);
    LK.cost(n, ld, xs, ys, box, cost.data(), vmin.data(), imin.data(), vmin2.data());
    lap.solve(n, ld, cost.data(), c4r.data(), imin.data(), vmin.data(), /*lazy_sqrt=*/box != nullptr);
    for (int i = 0; i < n; ++i) perm[idx[i]] = c4r[i] >= 0 ? idx[c4r[i]] : idx[i];
    if (margin >= 0) {
      seen.assign(n, 0);
      double gap = std::numeric_limits<double>::infinity();
      bool ok = true;
      for (int j = 0; j < n; ++j) {
        ok = ok && !seen[imin[j]] && vmin[j] == vmin[j];
        seen[imin[j]] = 1;
        gap = std::min(gap, __builtin_sqrt(vmin2[j]) - __builtin_sqrt(vmin[j]));  // entries are squared
      }
      margin = ok ? std::min(margin, gap) : -1.0;
    }
  }
  return margin;
}

// smallest eigenpair of a symmetric 4x4 matrix by cyclic Jacobi
void jacobi4(double A[4][4], double& eigmin, double q[4]) {
  double V[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0, diag = 0;
    for (int p = 0; p < 4; ++p) {
      diag += A[p][p] * A[p][p];
      for (int r = p + 1; r < 4; ++r) off += A[p][r] * A[p][r];
    }
    // off-diagonal norm below 1e-18 of the diagonal's: a further sweep cannot change a double (the former
    // absolute 1e-300 ran four more sweeps after that point, the last ones on denormals)
    if (off <= 1e-36 * diag) break;
    for (int p = 0; p < 4; ++p)
      for (int r = p + 1; r < 4; ++r) {
        if (fabs(A[p][r]) < 1e-300) continue;
        const double theta = (A[r][r] - A[p][p]) / (2 * A[p][r]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1));
        const double c = 1 / sqrt(t * t + 1), s = t * c;
        for (int k = 0; k < 4; ++k) {
          const double akp = A[k][p], akr = A[k][r];
          A[k][p] = c * akp - s * akr;
          A[k][r] = s * akp + c * akr;
        }
        for (int k = 0; k < 4; ++k) {
          const double apk = A[p][k], ark = A[r][k];
          A[p][k] = c * apk - s * ark;
          A[r][k] = s * apk + c * ark;
        }
        for (int k = 0; k < 4; ++k) {
          const double vkp = V[k][p], vkr = V[k][r];
          V[k][p] = c * vkp - s * vkr;
          V[k][r] = s * vkp + c * vkr;
        }
      }
  }
  int im = 0;
  for (int k = 1; k < 4; ++k)
    if (A[k][k] < A[im][im]) im = k;
  eigmin = A[im][im];
  for (int k = 0; k < 4; ++k) q[k] = V[k][im];
}

// Kearsley: distance after the optimal rotation of x2 onto x1 (both re-centred) and the matrix
double kearsley(int n, const double* x1, const double* x2, const int* perm, double R[9]) {
  double c1[3] = {0, 0, 0}, c2[3] = {0, 0, 0};
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < 3; ++k) {
      c1[k] += x1[3 * i + k];
      c2[k] += x2[3 * perm[i] + k];
    }
  for (int k = 0; k < 3; ++k) {
    c1[k] /= n;
    c2[k] /= n;
  }
  double Q[4][4] = {};
  for (int i = 0; i < n; ++i) {
    const double a[3] = {x1[3 * i] - c1[0], x1[3 * i + 1] - c1[1], x1[3 * i + 2] - c1[2]};
    const double b[3] = {x2[3 * perm[i]] - c2[0], x2[3 * perm[i] + 1] - c2[1], x2[3 * perm[i] + 2] - c2[2]};
    const double xm = a[0] - b[0], ym = a[1] - b[1], zm = a[2] - b[2];
    const double xp = a[0] + b[0], yp = a[1] + b[1], zp = a[2] + b[2];
    Q[0][0] += xm * xm + ym * ym + zm * zm;
    Q[0][1] += ym * zp - yp * zm;
    Q[0][2] += xp * zm - xm * zp;
    Q[0][3] += xm * yp - xp * ym;
    Q[1][1] += yp * yp + zp * zp + xm * xm;
    Q[1][2] += xm * ym - xp * yp;
    Q[1][3] += xm * zm - xp * zp;
    Q[2][2] += xp * xp + zp * zp + ym * ym;
    Q[2][3] += ym * zm - yp * zp;
    Q[3][3] += xp * xp + yp * yp + zm * zm;
  }
  for (int p = 0; p < 4; ++p)
    for (int r = 0; r < p; ++r) Q[p][r] = Q[r][p];
  double eig, q[4];
  jacobi4(Q, eig, q);
  if (eig < 0) eig = fabs(eig) < 1e-6 ? 0.0 : -eig;
  if (R) {
    const double nq = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    const double q0 = q[0] / nq, q1 = q[1] / nq, q2 = q[2] / nq, q3 = q[3] / nq;
    R[0] = 2 * (0.5 - q2 * q2 - q3 * q3); R[1] = 2 * (q1 * q2 - q0 * q3); R[2] = 2 * (q1 * q3 + q0 * q2);
    R[3] = 2 * (q1 * q2 + q0 * q3); R[4] = 2 * (0.5 - q1 * q1 - q3 * q3); R[5] = 2 * (q2 * q3 - q0 * q1);
    R[6] = 2 * (q1 * q3 - q0 * q2); R[7] = 2 * (q2 * q3 + q0 * q1); R[8] = 2 * (0.5 - q1 * q1 - q2 * q2);
  }
  return sqrt(eig);
}

// EulerM(a, b, y) = My Mb Ma (utils.py:447-460)
void euler_m(double a, double b, double y, double M[9]) {
  const double sa = sin(a), ca = cos(a), sb = sin(b), cb = cos(b), sy = sin(y), cy = cos(y);
  const double Ma[9] = {ca, -sa, 0, sa, ca, 0, 0, 0, 1};
  const double Mb[9] = {cb, 0, -sb, 0, 1, 0, sb, 0, cb};
  const double My[9] = {cy, -sy, 0, sy, cy, 0, 0, 0, 1};
  double T[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += My[3 * i + k] * Mb[3 * k + j];
      T[3 * i + j] = s;
    }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += T[3 * i + k] * Ma[3 * k + j];
      M[3 * i + j] = s;
    }
}

}  // namespace

// Test hook: selects the instruction set of the assignment kernels (see pick_lap_kernels; not thread safe: call
// while no refinement runs).  Returns the active one, or -1 when the requested one is not available on this CPU.
extern "C" int fo_host_lap_isa(int isa) {
  if (isa < 0) return g_lap_isa;
  LapKernels k;
  int picked = 0;
  if (!pick_lap_kernels(isa, &k, &picked)) return -1;
  LK = k;
  g_lap_isa = picked;
  return picked;
}

extern "C" void fo_host_refine_counters(int64_t out[3], int reset) {
  for (int k = 0; k < 3; ++k) {
    out[k] = (int64_t)__atomic_load_n(&g_counters[k], __ATOMIC_RELAXED);
    if (reset) __atomic_store_n(&g_counters[k], 0LL, __ATOMIC_RELAXED);
  }
}

namespace {
// pair_idx (nullable): the pairs to refine, results written at those indices; null = all of 0..npairs-1
int host_refine_periodic_impl(const fo_per_params* p, const int32_t* group_offsets, int64_t ngroups,
                              const int32_t* atom_idx, const double* posA, const double* posB,
                              const double* frac_idx, const int64_t* pair_idx, int64_t npairs, int niter,
                              int nthreads, double* dist, int32_t* perm_out, double* disp_out) {
  if (!p || !posA || !posB || !frac_idx || !dist || npairs < 0) return FO_ERR_INVALID;
  if (!groups_valid(group_offsets, ngroups, atom_idx, p->natoms)) return FO_ERR_INVALID;
  const int N = (int)p->natoms;
  const Groups G = {group_offsets, ngroups, atom_idx};
#ifdef _OPENMP
  int nt = nthreads > 0 ? nthreads : omp_get_num_procs();
  if ((int64_t)nt > npairs) nt = (int)std::max<int64_t>(npairs, 1);
#else
  const int nt = 1;
  (void)nthreads;
#endif
#pragma omp parallel num_threads(nt)
  {
    Lap lap;
    std::vector<double> cost, ys((size_t)3 * N), term((size_t)3 * N);
    std::vector<int> c4r, perm(N), save(N);
    std::vector<char> screen(ngroups);
#pragma omp for schedule(dynamic, 4)
    for (int64_t qq = 0; qq < npairs; ++qq) {
      const int64_t q = pair_idx ? pair_idx[qq] : qq;
      const double* x = posA + (size_t)q * N * 3;
      const double* y = posB + (size_t)q * N * 3;
      double disp[3];
      const double ibox[3] = {1.0 / p->box[0], 1.0 / p->box[1], 1.0 / p->box[2]};
      for (int k = 0; k < 3; ++k) disp[k] = frac_idx[3 * q + k] * p->box[k] / (double)p->nfspace;
      auto shift = [&]() {
        for (int i = 0; i < N; ++i)
          for (int k = 0; k < 3; ++k) ys[3 * i + k] = y[3 * i + k] - disp[k];
      };
      auto recentre = [&](const int* pm) {
        double* t = term.data();
        for (int i = 0; i < N; ++i)
          for (int k = 0; k < 3; ++k)
            t[k * N + i] = min_image(x[3 * i + k] - (y[3 * pm[i] + k] - disp[k]), p->box[k], ibox[k]);
        for (int k = 0; k < 3; ++k) disp[k] -= canon_sum(t + k * N, N) / N;
      };
      shift();
      std::fill(screen.begin(), screen.end(), 1);
      double margin = best_perm(G, N, x, ys.data(), p->box, lap, cost, c4r, save.data(), screen.data());
      double dref[3] = {disp[0], disp[1], disp[2]};  // displacement the margin belongs to
      perm = save;
      for (int it = 0; it < niter; ++it) {
        recentre(save.data());
        // The reference solves the assignment again and stops when it comes back unchanged
        // (periodicAlignment.py:66-75).  When the move since the last solve is provably too small to change
        // it (best_perm), the solve is skipped: same permutation, same displacement, bit for bit.
        const double dx = disp[0] - dref[0], dy = disp[1] - dref[1], dz = disp[2] - dref[2];
        if (margin > 0 && 2.0 * sqrt(dx * dx + dy * dy + dz * dz) + 1e-9 < margin) {
          count(2);
          break;
        }
        shift();
        margin = best_perm(G, N, x, ys.data(), p->box, lap, cost, c4r, perm.data(), screen.data());
        for (int k = 0; k < 3; ++k) dref[k] = disp[k];
        if (perm == save) break;
        save = perm;
      }
      recentre(perm.data());
      double* t = term.data();
      for (int i = 0; i < N; ++i) {
        double s = 0.0;
        for (int k = 0; k < 3; ++k) {
          // periodic(x) - periodic(y[perm] - disp), then the minimum image of the difference
          const double a = min_image(x[3 * i + k], p->box[k], ibox[k]);
          const double b = min_image(y[3 * perm[i] + k] - disp[k], p->box[k], ibox[k]);
          const double d = min_image(a - b, p->box[k], ibox[k]);
          s = k == 0 ? d * d : s + d * d;
        }
        t[i] = s;
      }
      dist[q] = sqrt(canon_sum(t, N));
      if (perm_out)
        for (int i = 0; i < N; ++i) perm_out[(size_t)q * N + i] = perm[i];
      if (disp_out)
        for (int k = 0; k < 3; ++k) disp_out[3 * q + k] = disp[k];
    }
  }
  return FO_OK;
}
}  // namespace

extern "C" int fo_host_refine_periodic(const fo_per_params* p, const int32_t* group_offsets, int64_t ngroups,
                                       const int32_t* atom_idx, const double* posA, const double* posB,
                                       const double* frac_idx, int64_t npairs, int niter, int nthreads,
                                       double* dist, int32_t* perm_out, double* disp_out) {
  return host_refine_periodic_impl(p, group_offsets, ngroups, atom_idx, posA, posB, frac_idx, nullptr, npairs,
                                   niter, nthreads, dist, perm_out, disp_out);
}

extern "C" int fo_host_refine_periodic_subset(const fo_per_params* p, const int32_t* group_offsets,
                                              int64_t ngroups, const int32_t* atom_idx, const double* posA,
                                              const double* posB, const double* frac_idx,
                                              const int64_t* pair_idx, int64_t nidx, int niter, int nthreads,
                                              double* dist, int32_t* perm_out, double* disp_out) {
  if (nidx > 0 && !pair_idx) return FO_ERR_INVALID;
  return host_refine_periodic_impl(p, group_offsets, ngroups, atom_idx, posA, posB, frac_idx, pair_idx, nidx,
                                   niter, nthreads, dist, perm_out, disp_out);
}

namespace {
// perm_hint [P, norient, N] / hint_ok [P, norient] (nullable): permutations the device screening proved optimal
// (fo_assign.cu: sph_assign_kernel); where hint_ok is set the LAP is skipped.
int host_refine_spherical_impl(const double* posA, const double* posB, int64_t npairs, int64_t natoms,
                               const int32_t* group_offsets, int64_t ngroups, const int32_t* atom_idx,
                               const double* euler, int norient, const int32_t* perm_hint,
                               const int32_t* hint_ok, int nthreads, double* dist, int32_t* orient_out,
                               int32_t* perm_out, double* rmat_out, const int64_t* pair_idx = nullptr,
                               const double* pre_dist = nullptr, const double* pre_rot = nullptr) {
  // pair_idx: npairs indices into the arrays (a subset of a larger batch), or null for pairs 0 .. npairs - 1
  // pre_dist [P, norient] / pre_rot [P, norient, 9]: Kearsley distance and rotation the device already computed for
  // the orientations whose hint_ok is set (sph_assign_kernel): those orientations cost nothing here
  if (!posA || !posB || !euler || !dist || npairs < 0 || natoms < 1 || norient < 1 || norient > 2)
    return FO_ERR_INVALID;
  if (!groups_valid(group_offsets, ngroups, atom_idx, natoms)) return FO_ERR_INVALID;
  if ((perm_hint == nullptr) != (hint_ok == nullptr)) return FO_ERR_INVALID;
  const int N = (int)natoms;
  const Groups G = {group_offsets, ngroups, atom_idx};
#ifdef _OPENMP
  int nt = nthreads > 0 ? nthreads : omp_get_num_procs();
  if ((int64_t)nt > npairs) nt = (int)std::max<int64_t>(npairs, 1);
#else
  const int nt = 1;
  (void)nthreads;
#endif
#pragma omp parallel num_threads(nt)
  {
    Lap lap;
    std::vector<double> cost, xr((size_t)3 * N);
    std::vector<int> c4r, perm(N), bestperm(N);
#pragma omp for schedule(dynamic, 8)
    for (int64_t qi = 0; qi < npairs; ++qi) {
      const int64_t q = pair_idx ? pair_idx[qi] : qi;
      const double* x1 = posA + (size_t)q * N * 3;
      const double* x2 = posB + (size_t)q * N * 3;
      double best = std::numeric_limits<double>::infinity(), bestR[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
      int besto = 0;
      for (int o = 0; o < norient; ++o) {
        if (pre_dist && hint_ok && hint_ok[q * norient + o]) {  // settled and fitted on the device
          const double d = pre_dist[q * norient + o];
          count(1);
          if (d < best) {
            best = d;
            besto = o;
            const int32_t* h = perm_hint + ((size_t)q * norient + o) * N;
            for (int i = 0; i < N; ++i) bestperm[i] = h[i];
            memcpy(bestR, pre_rot + ((size_t)q * norient + o) * 9, sizeof(bestR));
          }
          continue;
        }
        const double* e = euler + ((size_t)q * norient + o) * 3;
        double M[9], R[9];
        euler_m(e[0], e[1], e[2], M);
        const double sg = o ? -1.0 : 1.0;  // orientation 1: inverted structure -X2
        for (int i = 0; i < N; ++i)
          for (int j = 0; j < 3; ++j) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += sg * x2[3 * i + k] * M[3 * k + j];  // (X2 . M)
            xr[3 * i + j] = s;
          }
        if (hint_ok && hint_ok[q * norient + o]) {
          const int32_t* h = perm_hint + ((size_t)q * norient + o) * N;
          for (int i = 0; i < N; ++i) perm[i] = h[i];
          count(1);
        } else {
          best_perm(G, N, x1, xr.data(), nullptr, lap, cost, c4r, perm.data());
        }
        const double d = kearsley(N, x1, xr.data(), perm.data(), R);
        if (d < best) {
          best = d;
          besto = o;
          bestperm = perm;
          memcpy(bestR, R, sizeof(R));
        }
      }
      dist[q] = best;
      if (orient_out) orient_out[q] = besto;
      if (perm_out)
        for (int i = 0; i < N; ++i) perm_out[(size_t)q * N + i] = bestperm[i];
      if (rmat_out) memcpy(rmat_out + 9 * q, bestR, sizeof(bestR));
    }
  }
  return FO_OK;
}
}  // namespace

extern "C" int fo_host_refine_spherical(const double* posA, const double* posB, int64_t npairs, int64_t natoms,
                                        const int32_t* group_offsets, int64_t ngroups, const int32_t* atom_idx,
                                        const double* euler, int norient, int nthreads, double* dist,
                                        int32_t* orient_out, int32_t* perm_out, double* rmat_out) {
  return host_refine_spherical_impl(posA, posB, npairs, natoms, group_offsets, ngroups, atom_idx, euler, norient,
                                    nullptr, nullptr, nthreads, dist, orient_out, perm_out, rmat_out);
}

extern "C" int fo_host_refine_spherical_hint(const double* posA, const double* posB, int64_t npairs,
                                             int64_t natoms, const int32_t* group_offsets, int64_t ngroups,
                                             const int32_t* atom_idx, const double* euler, int norient,
                                             const int32_t* perm_hint, const int32_t* hint_ok, int nthreads,
                                             double* dist, int32_t* orient_out, int32_t* perm_out,
                                             double* rmat_out) {
  return host_refine_spherical_impl(posA, posB, npairs, natoms, group_offsets, ngroups, atom_idx, euler, norient,
                                    perm_hint, hint_ok, nthreads, dist, orient_out, perm_out, rmat_out);
}

// the pairs pair_idx[0 .. nidx) of a batch (the ones the device did not settle)
int fo_host_refine_spherical_subset(const double* posA, const double* posB, int64_t natoms,
                                    const int32_t* group_offsets, int64_t ngroups, const int32_t* atom_idx,
                                    const double* euler, int norient, const int32_t* perm_hint, const int32_t* hint_ok,
                                    const int64_t* pair_idx, int64_t nidx, int nthreads, double* dist,
                                    int32_t* orient_out, int32_t* perm_out, double* rmat_out, const double* pre_dist,
                                    const double* pre_rot) {
  if (nidx > 0 && !pair_idx) return FO_ERR_INVALID;
  return host_refine_spherical_impl(posA, posB, nidx, natoms, group_offsets, ngroups, atom_idx, euler, norient,
                                    perm_hint, hint_ok, nthreads, dist, orient_out, perm_out, rmat_out, pair_idx,
                                    pre_dist, pre_rot);
}

// ---------------------------------------------------------------------------------------------------------
// The two host steps on their own, for the single-pair drop-in classes (Hungarian / findrotation of
// utils.py:82-253): same code the pools above run per pair.

extern "C" int fo_host_best_permutation(const double* posA, const double* posB, int64_t natoms,
                                        const int32_t* group_offsets, int64_t ngroups, const int32_t* atom_idx,
                                        const double* box, int32_t* perm_out) {
  if (!posA || !posB || !perm_out) return FO_ERR_INVALID;
  if (!groups_valid(group_offsets, ngroups, atom_idx, natoms)) return FO_ERR_INVALID;
  const Groups G = {group_offsets, ngroups, atom_idx};
  Lap lap;
  std::vector<double> cost;
  std::vector<int> c4r, perm((size_t)natoms);
  best_perm(G, (int)natoms, posA, posB, box, lap, cost, c4r, perm.data());
  for (int64_t i = 0; i < natoms; ++i) perm_out[i] = perm[i];
  return FO_OK;
}

extern "C" int fo_host_kearsley(const double* x1, const double* x2, int64_t natoms, double* dist, double* rmat) {
  if (!x1 || !x2 || !dist || natoms < 1) return FO_ERR_INVALID;
  std::vector<int> id((size_t)natoms);
  for (int64_t i = 0; i < natoms; ++i) id[i] = (int)i;
  *dist = kearsley((int)natoms, x1, x2, id.data(), rmat);
  return FO_OK;
}
