// Internal declarations shared by the translation units of libfastoverlap_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <algorithm>
#include <string>
#include <map>
#include <string>
#include <vector>

#include "../../include/fastoverlap_b200.h"

struct fo_devbuf {
  void* ptr = nullptr;
  size_t bytes = 0;
};

// Cached Wigner-d table for one bandwidth (device resident).
struct fo_wigner_cache {
  int64_t Jmax = -1;
  double* d_table = nullptr;   // dense Dt[m2][m1][l][k], see fo_spherical.cu
  double* d_packed = nullptr;  // per-chunk shell-ordered slices for sph_isoft3_kernel
  int packed_kc = 0;           // beta planes per chunk of d_packed (0: none)
  double* d_packed4 = nullptr; // slices of four planes (sph_isoft4_kernel<4, ...>, odd Jmax <= 15)
  size_t bytes = 0;
  bool kmajor = false;         // large bandwidths: plane-major layout DtK[k][m2][m1][l]
};

struct fo_prof_rec {
  int kind;
  cudaEvent_t e0, e1;
};

struct fo_ctx {
  int device = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;  // stream in use (own or external)
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // 0,1 H2D done; 2,3 inputs consumed; 4,5 D2H done
  cudaDeviceProp prop;
  std::string err;
  int64_t launches = 0;

  // permutation groups (host + device copies)
  std::vector<int32_t> h_goff;  // [ngroups+1]
  std::vector<int32_t> h_gidx;  // [natoms_in_groups]
  int32_t* d_goff = nullptr;
  int32_t* d_gidx = nullptr;
  int64_t perm_natoms = 0;  // natoms the perm was declared for (0 = unset -> one group of all)
  int64_t gid_natoms = -1;  // natoms the cached per-atom group-id table (FO_SCR_GID) was built for; -1 = stale

  // growable device scratch (named slots)
  fo_devbuf scratch[14];
  // pinned host staging (named slots)
  fo_devbuf pinned[8];

  fo_wigner_cache wig;
  // recurrence table of the continuous rotation refinement (fo_refine.cu), cached per bandwidth
  fo_devbuf refine_tab;
  int refine_L = -1;

  // testing hook: force the generic (any-size) kernels instead of the shared-memory fast paths
  bool force_generic = false;
  // A/B hook: 0 = default iSOFT kernel, 3 = sph_isoft3_kernel (stage A -> shared memory -> stage B)
  int isoft_variant = 0;
  // independent pairs at n = 9: fused structure factors + cross-spectrum (no bank); 0 = bank path (A/B, tests)
  bool pairs_fused = true;
  // tuning / A-B hooks set through fo_set_option (no environment variables): per_sf_scalar, per_sf_padded,
  // per_sf_tile_atoms, per_sf_syncthreads, per_xf_generic, per_chunk_mb, sph_chunk_mb
  std::map<std::string, int64_t> tune;
  int64_t opt(const char* name, int64_t dflt = 0) const {
    const auto it = tune.find(name);
    return it == tune.end() ? dflt : it->second;
  }
  // clusters with at least this many atoms use the tensor-core GEMM form of the direct coefficients
  int64_t direct_gemm_min = 64;

  // per-kernel event timing (fo_profile_begin/end)
  bool profiling = false;
  std::vector<fo_prof_rec> prof;
};

enum {
  FO_SCR_POSA = 0,
  FO_SCR_POSB = 1,
  FO_SCR_BANK = 2,
  FO_SCR_OUT = 3,
  FO_SCR_GRID = 4,
  FO_SCR_WORK = 5,
  FO_SCR_COEF = 6,
  FO_SCR_MISC = 7,
  FO_SCR_IPK = 8,
  FO_SCR_DBG = 9,
  FO_SCR_PEAKS = 10,
  FO_SCR_FULL = 11,
  FO_SCR_GID = 12,
};

struct fo_bank {
  int kind = 0;  // 1 = periodic structure factors, 2 = spherical harmonic coefficients
  int64_t nstruct = 0;
  int64_t ngroups = 0;
  int64_t per_struct_elems = 0;  // double2 elements per structure
  // periodic
  int64_t nwave = 0;
  // spherical
  int64_t nmax = 0, Jmax = 0, natoms = 0;
  double sigma = 0, harmscale = 0;
  double* d_data = nullptr;
};

int fo_fail(fo_ctx* ctx, int code, const char* fmt, ...);
int fo_scratch(fo_ctx* ctx, int slot, size_t bytes, void** out);
int fo_pinned(fo_ctx* ctx, int slot, size_t bytes, void** out);
// true when `p` is page-locked (cudaMallocHost / cudaHostRegister): it can be DMA'd directly
bool fo_is_pinned(const void* p);
// parallel host memcpy (staging of pageable buffers into the pinned ring)
void fo_host_copy(void* dst, const void* src, size_t bytes, int nthreads = 0);
// Chunk boundaries of the double-buffered host-buffer pipelines: [0, s1, s2, ..., npairs], every
// chunk <= chunk pairs.  With more than one chunk the sizes ramp up geometrically (chunk / 8, x2, x2, ...):
// the first H2D copy is the only one no kernel overlaps, and the kernels of chunk c must last at least
// as long as the copy of chunk c + 1 (BLJ256: 0.25 us / pair of PCIe against 0.65 us / pair of compute).
inline std::vector<int64_t> fo_chunk_starts(int64_t npairs, int64_t chunk) {
  // Chunk boundaries of the host pipelines: short chunks at BOTH ends -- the H2D copy of the first chunk and the
  // D2H + delivery of the last one are the only transfers nothing can hide -- ramping by factors of two to the
  // full chunk size in between.
  std::vector<int64_t> s(1, 0);
  if (npairs <= chunk || chunk < 64) {
    for (int64_t p0 = 0; p0 < npairs;) {
      p0 = std::min(p0 + chunk, npairs);
      s.push_back(p0);
    }
    return s;
  }
  const int64_t ramp[3] = {chunk / 8, chunk / 4, chunk / 2};
  const int64_t ends = 2 * (ramp[0] + ramp[1] + ramp[2]);
  std::vector<int64_t> sizes;
  if (npairs >= ends + chunk / 2) {
    for (int i = 0; i < 3; ++i) sizes.push_back(ramp[i]);
    const int64_t mid = npairs - ends;
    const int64_t nmid = (mid + chunk - 1) / chunk;
    for (int64_t i = 0; i < nmid; ++i) sizes.push_back(mid / nmid + (i < mid % nmid ? 1 : 0));
    for (int i = 2; i >= 0; --i) sizes.push_back(ramp[i]);
  } else {  // a few chunks only: ramp up as far as it goes
    int64_t size = chunk / 8, left = npairs;
    while (left > 0) {
      const int64_t n = std::min(size, left);
      sizes.push_back(n);
      left -= n;
      size = std::min(2 * size, chunk);
    }
  }
  int64_t p0 = 0;
  for (int64_t n : sizes) {
    p0 += n;
    s.push_back(p0);
  }
  return s;
}

// host pool on a subset of a batch of cluster pairs (fo_host.cu; internal)
int fo_host_refine_spherical_subset(const double* posA, const double* posB, int64_t natoms,
                                    const int32_t* group_offsets, int64_t ngroups, const int32_t* atom_idx,
                                    const double* euler, int norient, const int32_t* perm_hint, const int32_t* hint_ok,
                                    const int64_t* pair_idx, int64_t nidx, int nthreads, double* dist,
                                    int32_t* orient_out, int32_t* perm_out, double* rmat_out,
                                    const double* pre_dist = nullptr, const double* pre_rot = nullptr);

// make sure a permutation (at least the trivial one) exists for natoms atoms
int fo_ensure_perm(fo_ctx* ctx, int64_t natoms);

// top-k peak extraction on device grids (fo_peaks.cu); grids are updated in place
int fo_peaks_run_dev(fo_ctx* ctx, double* d_grids, int64_t P, const int64_t shape[3], int64_t npeaks,
                     int64_t width, double* d_peaks, double* d_amp, double* d_mean, double* d_alpha,
                     int32_t* d_nfound);
int fo_peaks_outputs(fo_ctx* ctx, int64_t np, int64_t npeaks, double** pk, double** amp, double** mean,
                     double** alpha, int32_t** nf);
int fo_peaks_copy_out(fo_ctx* ctx, int64_t p0, int64_t np, int64_t npeaks, const double* pk, const double* amp,
                      const double* mean, const double* alpha, const int32_t* nf, double* peaks,
                      double* amplitude, double* meanv, double* alphav, int32_t* nfound);

// continuous rotation refinement (fo_refine.cu): one CTA per (pair, orientation) on the Ihalf layout;
// d_start is [np*norient][3] fractional grid indices (from_frac) or Euler angles
int fo_refine_run_dev(fo_ctx* ctx, const void* d_Ihalf, int64_t np, int L, int norient, const double* d_start,
                      int from_frac, double* d_euler, double* d_overlap, int* d_iters);

// device-side screening of the permutational assignment (fo_assign.cu).  Periodic: d_flag [np] = 0 when the pair
// was settled on the device (dist / disp / perm written), non-zero when the host LAP pool must take it.
// Clusters: d_perm [np, norient, natoms], d_ok [np, norient] = 1 where the permutation is the proven optimum.
int fo_per_assign_run_dev(fo_ctx* ctx, const fo_per_params* p, const double* d_posA, const double* d_posB,
                          const double* d_frac, int64_t np, int niter, double* d_dist, double* d_disp,
                          void* d_perm, int32_t* d_flag, int perm_elt = 4);  // perm_elt: bytes per index (4, 2, 1)
int fo_sph_assign_run_dev(fo_ctx* ctx, const double* d_posA, const double* d_posB, const double* d_frac,
                          int64_t np, int64_t natoms, int L, int norient, int32_t* d_perm, int32_t* d_ok,
                          double* d_kdist = nullptr, double* d_krot = nullptr);  // Kearsley fit of settled assignments

int fo_refine_eval_dev(fo_ctx* ctx, const void* d_Ihalf, int64_t np, int L, const double* d_euler, double* d_value,
                       double* d_grad, double* d_hess);

// RAII bracket: records an event pair around a kernel launch when profiling is on
struct fo_prof_scope {
  fo_ctx* ctx;
  cudaEvent_t e1 = nullptr;
  fo_prof_scope(fo_ctx* c, int kind) : ctx(c) {
    if (!c->profiling) return;
    cudaEvent_t e0;
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) {
      e1 = nullptr;
      return;
    }
    cudaEventRecord(e0, c->stream);
    c->prof.push_back({kind, e0, e1});
  }
  ~fo_prof_scope() {
    if (e1) cudaEventRecord(e1, ctx->stream);
  }
};

#define FO_CUDA(ctx, call)                                                              \
  do {                                                                                  \
    cudaError_t _e = (call);                                                            \
    if (_e != cudaSuccess)                                                              \
      return fo_fail((ctx), FO_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__,     \
                     __LINE__, cudaGetErrorString(_e));                                 \
  } while (0)

#define FO_CHECK(expr)            \
  do {                            \
    int _rc = (expr);             \
    if (_rc != FO_OK) return _rc; \
  } while (0)

#define FO_LAUNCH_CHECK(ctx)                                                            \
  do {                                                                                  \
    (ctx)->launches++;                                                                  \
    cudaError_t _e = cudaGetLastError();                                                \
    if (_e != cudaSuccess)                                                              \
      return fo_fail((ctx), FO_ERR_CUDA, "kernel launch failed at %s:%d: %s", __FILE__, \
                     __LINE__, cudaGetErrorString(_e));                                 \
  } while (0)
