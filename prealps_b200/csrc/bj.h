// bj.h -- shared layout of the block-Jacobi factor between bj_factor.cu (numeric
// supernodal Cholesky on the device) and bj_solve.cu (level-scheduled sweeps).
//
// Factor layout in HBM (see DESIGN.md "Block-Jacobi"):
//   every supernode s of the forest (all local diagonal blocks together) has a dense
//   h x w trapezoid  M_s = [ L_ss^{-1} ; L_bs L_ss^{-1} ]  so that both sweeps are pure
//   streaming products without any intra-supernode dependency:
//     forward   [y_s ; u_s] = M_s b_s          (u_s = contribution to the ancestors)
//     backward  x_s = M_s^T [y_s ; -x_below]
//   M_s is cut into 32-row slices ("panels") of klen steps (klen a multiple of 4); the rows below the diagonal block
//   are stored negated:
//     fwd panel p of s: rows [32p, 32p+32) of M_s, steps k = columns [0, min(w, 32p+32)) of M_s; produces y_s and -u_s
//     bwd panel q of s: columns [32q, 32q+32) of M_s = k-blocks [8q, 8q+8) of every fwd panel p >= q, walked tile by
//                       tile (32 rows x 32 columns = 8 KB contiguous); steps k = rows 32q.. of M_s
//   A panel is a sequence of k-blocks (4 steps, 128 doubles = 1 KB) stored in the A-fragment order of
//   mma.sync.m8n8k4.f64: data[kb*128 + lane*4 + rg] = panel(8*rg + lane/4, 4*kb + lane%4).  A warp streams a
//   panel with perfectly coalesced 1 KB reads (32 B per lane) and multiplies it with the t-wide input rows
//   (B fragment: one shared-memory double per lane and k-block) on the FP64 tensor cores.  The backward sweep reads the
//   same k-blocks with a permuted lane -> address map that yields the A fragments of M_s^T directly (bj_solve.cu:
//   bwd_lane_offset), so ONE copy of the factor serves both sweeps.
//   The panels of a supernode are contiguous, slice after slice: slice p starts panel_cum(w, p) doubles after the
//   supernode's first one, so the backward sweep needs no per-slice offset table.
//   When memory allows (bj_factor.cu: tcopy) the factor also keeps M_s^T cut into its own 32-row slices (bwd_data): the
//   backward sweep then streams contiguous panels like the forward one, ~10 % faster than the strided 8 KB tiles of the
//   single copy on B200 (same bytes; measured, profiles/r02_single_copy_factor.md).
#pragma once
#include <cstdint>
#include <vector>

#include "common.cuh"

namespace pcu {

// doubles stored in front of slice p of a supernode of width w: slices p' < p hold min(W4, 8p'+8) k-blocks of 128
// doubles each, W4 = ceil(w / 4)
__host__ __device__ inline long long panel_cum(int w, int p) {
  const int W4 = (w + 3) >> 2, pt = W4 >> 3, m = p < pt ? p : pt;
  return 128ll * (4ll * m * (m + 1) + (long long)(p > pt ? p - pt : 0) * W4);
}

struct FwdPanel {      // one 32-row slice of M_s
  long long off;       // offset (doubles) into fwd_data
  long long uoff;      // first row of this supernode's slot in the update buffer U
  int klen;            // number of k steps (even)
  int c0;              // first column of the supernode (forest index): input rows Wk[c0 .. c0+klen)
  int row0;            // first row of the slice inside the supernode (multiple of 32)
  int w, h;            // supernode width / height
  int pad_;
};

struct BwdPanel {      // one 32-row slice of M_s^T = 32 columns of M_s
  long long off;       // one copy: offset of the SUPERNODE's first slice in fwd_data (the slices follow: panel_cum);
                       // with the transposed copy: offset of the slice in bwd_data
  long long rows_off;  // offset into rows (forest row index of each supernode row)
  int klen;            // number of k steps covering supernode rows [k0, k0+klen) (zero rows past h): whole 32-row tiles
                       // (one copy) or whole k-blocks of 4 (transposed copy)
  int k0;              // = 32 q: first supernode row that contributes = first column of the slice
  int c0;              // first column of the supernode
  int w, h;
  int pad_;
};

constexpr int kChunkMinKB = 128;            // shortest slice (k-blocks of 4 steps) a CTA gets when a long panel is cut across CTAs
constexpr int kWarpSlots = 148 * 2 * 8;     // resident warps of the sweep kernel (2 CTAs of 8 warps per SM)

struct WorkUnit {      // one CTA of the sweep kernels
  int first;           // first panel
  int count;           // 1..8 panels (one per warp), or 1 panel split over all warps when split != 0
  int split;           // 0: one warp per panel; 1: the 8 warps share one panel; 2: ... and only its k-blocks [kb0, kb1)
  int kb0, kb1;        // split == 2: this CTA's slice of the panel
  int chunk, nchunks;  // split == 2: position of the slice; the CTA that finishes last adds the slices in order
  int slot;            // split == 2: first scratch slot of the panel (one slot = 32 x T doubles per slice); counter index in cidx
  int cidx;
  int pad_;
};

}  // namespace pcu

struct pcu_bj {
  pcu_ctx* ctx = nullptr;
  int n = 0;           // total rows over all local blocks
  int nblk = 0;
  int nsuper = 0, nlevels = 0;
  long long nu = 0;    // rows of the update buffer = sum (h - w)
  double stat[16] = {0};
  // device: factor
  double* fwd_data = nullptr;       // the panels (one copy, both sweeps)
  double* bwd_data = nullptr;       // optional transposed copy (null: the backward sweep reads fwd_data tile by tile)
  long long fwd_doubles = 0, bwd_doubles = 0;
  pcu::FwdPanel* fwd_panels = nullptr;
  pcu::BwdPanel* bwd_panels = nullptr;
  pcu::WorkUnit* fwd_units = nullptr;
  pcu::WorkUnit* bwd_units = nullptr;
  std::vector<int> fwd_unit_ptr, bwd_unit_ptr;   // per level, nlevels+1
  std::vector<double> fwd_lvl_bytes, bwd_lvl_bytes;    // panel bytes per level (profiling)
  // device: assembly of the forward right-hand side
  int* perm = nullptr;             // perm[forest col] = local row of the m x t block
  int* rows = nullptr;             // forest row index of every supernode row (gather index of the backward sweep)
  int* lvl_cols = nullptr;         // forest columns sorted by level
  std::vector<int> lvl_col_ptr;    // per level, nlevels+1
  std::vector<char> lvl_long_lists;  // per level: its columns gather > 16 update rows on average (assemble_kernel<T, 16>)
  long long* gl_ptr = nullptr;     // per forest column: range into gl_idx
  long long* gl_idx = nullptr;     // rows of U that must be subtracted from that column
  // work vectors (sized for cap_t columns)
  int cap_t = 0;
  double* Wk = nullptr;
  double* Y = nullptr;
  double* U = nullptr;
  double* Xp = nullptr;            // solution in forest (permuted) order, read by the descendants
  // inter-CTA split-K: partial results of the slices of one panel and one arrival counter per panel
  double* scratch = nullptr;
  int* counters = nullptr;
  int scratch_slots = 0, ncounters = 0;
};
