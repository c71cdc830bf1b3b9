// river_kernel.cuh — musk_classic routing over the node/reach DAG as a topological-level wavefront.
//
// Replaces `c_musk_classic.pyx` (`Pick_Inflow_V1`, `Update_Discharge_V1`, `Calc_Discharge_V1`, `Calc_Outflow_V1`,
// `Pass_Outflow_V1`; ref: hydpy/models/musk/musk_model.py:19-187, 809-848; SegmentModel.run
// hydpy/core/modeltools.py:3449-3461) and the node summation of the reference's whole-period mode
// (ref: hydpy/core/threadingtools.py:396-446).
//
// All series are device rows [row][nt] (time contiguous):  land rows (outlets.q of the land elements), node rows,
// reach inflow/outflow rows, external rows.
//
//   reach_diag_kernel : one warp per (reach, time chunk).  The window is cut into chunks of HB_RIVER_CHUNK steps and
//                     the (level, chunk) grid is swept along its anti-diagonals: launch d works on every reach of level l
//                     at chunk d-l, which only needs chunk d-l of the levels below — finished by launch d-1.  The
//                     sequential depth drops from levels x nt ticks to (levels + nt/chunk) short launches and every
//                     launch has nt/chunk levels' worth of reaches in flight.
//                     Inside a warp lane i owns segment i and the chunk runs as a systolic pipeline over "ticks": at
//                     tick τ lane i computes time τ-i, receiving new[i](t) from lane i-1 by shuffle.  The recurrence
//                     new[i+1] = c0*new[i] + c1*old[i] + c2*old[i+1]  is evaluated with the reference's operation order
//                     and no FMA contraction ⇒ bit-identical to the CPU path.  The warp also forms the sums of its inlet
//                     nodes (entries added in the listed order, 0.0 + a + b + …) and, as the owner of a node, stores
//                     the node row.
//   node_sum_kernel : one warp per node, lanes over time — for the nodes without a downstream reach (outlets).
#pragma once
#include <cstdint>

#include "../../include/hydpy_b200.h"

namespace hb {

struct RiverDev {
  int64_t n_node, n_reach, nt;
  const int64_t* entry_ptr;
  const int32_t* entry_src;
  const int32_t* node_mode;
  const int32_t* node_ext;
  const int64_t* inlet_ptr;
  const int32_t* inlet_node;
  const int64_t* seg_ptr;
  const double* coef;       // [3][n_reach]
  const double* land_rows;  // [n_elem][nt]
  const double* ext_rows;   // [n_ext][nt]
  double* node_rows;        // [n_node][nt]
  double* reach_in;         // [n_reach][nt]
  double* reach_out;        // [n_reach][nt]
  double* dis[2];           // chunked sweep: chunk c reads dis[(par0 + c) & 1] and writes the other one
  int par0;
  const int32_t* reach_level;  // [n_reach]
  const int32_t* node_owner;   // [n_node] reach that stores the node row (first exit), -1: terminal node
};

#define HB_RIVER_CHUNK 32
#define HB_BULK_SMAX 32      // reaches with more segments always take the systolic path
#define HB_BULK_WARPS 4      // warps per block of reach_bulk_kernel
#define HB_BULK_SMEM (HB_BULK_WARPS * (32 * 33 + (HB_BULK_SMAX + 1) * 32) * 8)

// qt [nt][e_pad]  →  land rows [n_elem][nt]   (32x32 shared-memory tile transpose, both sides coalesced)
__global__ void __launch_bounds__(256) transpose_qt_kernel(const double* __restrict__ qt, double* __restrict__ rows,
                                                           int64_t nt, int64_t n_elem, int64_t e_pad) {
  __shared__ double tile[32][33];
  const int64_t e0 = (int64_t)blockIdx.x * 32, t0 = (int64_t)blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int j = ty; j < 32; j += 8) {
    const int64_t t = t0 + j, e = e0 + tx;
    tile[j][tx] = (t < nt && e < n_elem) ? qt[t * e_pad + e] : 0.0;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int64_t e = e0 + j, t = t0 + tx;
    if (e < n_elem && t < nt) rows[e * nt + t] = tile[tx][j];
  }
}

// ref: hydpy/core/threadingtools.py:396-415 `_update_node_series`: series = 0.0; series += entry series (in order)
__global__ void __launch_bounds__(256) node_sum_kernel(const RiverDev d, const int32_t* __restrict__ nodes, int n) {
  const int warp = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (warp >= n) return;
  const int64_t node = nodes[warp];
  const int64_t i0 = d.entry_ptr[node], i1 = d.entry_ptr[node + 1];
  double* row = d.node_rows + node * d.nt;
  for (int64_t t = lane; t < d.nt; t += 32) {
    double acc = 0.0;
    for (int64_t i = i0; i < i1; ++i) {
      const int32_t s = d.entry_src[i];
      const double* src = (s >= 0)             ? d.land_rows + (int64_t)s * d.nt
                          : (s > HB_ENTRY_EXT) ? d.reach_out + (int64_t)(-1 - s) * d.nt
                                               : d.ext_rows + (int64_t)(HB_ENTRY_EXT - s) * d.nt;
      acc += src[t];
    }
    row[t] = acc;
  }
}

// ref: threadingtools.py:396-415 node series = 0.0 + entries in order
__device__ __forceinline__ double node_sum_at(const RiverDev& d, int64_t node, int64_t t) {
  double acc = 0.0;
  const int64_t i0 = d.entry_ptr[node], i1 = d.entry_ptr[node + 1];
  for (int64_t i = i0; i < i1; ++i) {
    const int32_t s = d.entry_src[i];
    const double* src = (s >= 0)             ? d.land_rows + (int64_t)s * d.nt
                        : (s > HB_ENTRY_EXT) ? d.reach_out + (int64_t)(-1 - s) * d.nt
                                             : d.ext_rows + (int64_t)(HB_ENTRY_EXT - s) * d.nt;
    acc += src[t];
  }
  return acc;
}

// bulk != 0: reaches with 1..HB_BULK_SMAX segments only get their inflow here; reach_bulk_kernel runs their recurrence
__global__ void __launch_bounds__(128) reach_diag_kernel(const RiverDev d, const int32_t* __restrict__ reaches, int n,
                                                         int diag, int n_chunks, int bulk) {
  const int warp = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (warp >= n) return;
  const unsigned full = 0xffffffffu;
  const int64_t r = reaches[warp];
  const int c = diag - d.reach_level[r];
  if (c < 0 || c >= n_chunks) return;
  const int64_t nt = d.nt;
  const int64_t ta = (int64_t)c * HB_RIVER_CHUNK;
  const int64_t tb = (ta + HB_RIVER_CHUNK < nt) ? ta + HB_RIVER_CHUNK : nt;
  const int len = (int)(tb - ta);
  const double* dis_old = d.dis[(d.par0 + c) & 1];
  double* dis_new = d.dis[(d.par0 + c + 1) & 1];
  const int64_t g0 = d.seg_ptr[r];
  const int S = (int)(d.seg_ptr[r + 1] - g0 - 1);
  const double c0 = d.coef[0 * d.n_reach + r], c1 = d.coef[1 * d.n_reach + r], c2 = d.coef[2 * d.n_reach + r];
  const int64_t j0 = d.inlet_ptr[r], j1 = d.inlet_ptr[r + 1];
  double* rin = d.reach_in + r * nt;
  double* rout = d.reach_out + r * nt;

  // ---- inflow of the chunk (lane = time step): node sums, node rows of owned nodes, Pick_Inflow_V1 ----------------------
  double x_own = 0.0;   // inflow at time ta+lane, kept in the register for the pipeline below
  if (lane < len) {
    const int64_t t = ta + lane;
    double inflow = 0.0;
    for (int64_t j = j0; j < j1; ++j) {
      const int32_t node = d.inlet_node[j];
      const double sim = node_sum_at(d, node, t);
      if (d.node_owner[node] == (int32_t)r) d.node_rows[(int64_t)node * nt + t] = sim;
      const int mode = d.node_mode[node];
      double v = sim;
      if (mode != HB_NODE_NEWSIM) {
        v = d.ext_rows[(int64_t)d.node_ext[node] * nt + t];
        if (mode == HB_NODE_OBSNEWSIM && isnan(v)) v = sim;
      }
      inflow += 0.0 + v;   // ref: musk_model.py:26-31 Pick_Inflow_V1
    }
    rin[t] = inflow;
    if (S == 0) rout[t] = inflow;  // ref: musk_model.py:95-98, 831-835 — zero segments: outflow = discharge[0]
    x_own = inflow;
  }
  const double x_last = __shfl_sync(full, x_own, len - 1);
  if (lane == 0) dis_new[g0] = x_last;
  if (S == 0 || (bulk && S <= HB_BULK_SMAX)) return;

  // ---- systolic pipeline over groups of 32 segments --------------------------------------------------------------------
  for (int cbase = 0; cbase < S; cbase += 32) {
    const int nseg = (S - cbase < 32) ? (S - cbase) : 32;
    double prev_in = 0.0, prev_out = 0.0, out = 0.0;
    if (lane < nseg) {
      prev_in = dis_old[g0 + cbase + lane];
      prev_out = dis_old[g0 + cbase + lane + 1];
    }
    double xb = x_own;                               // input of this segment group at time ta+lane
    double yb = 0.0;                                 // its output at time ta+lane (collected below)
    const int nticks = len + nseg - 1;
    for (int tau = 0; tau < nticks; ++tau) {
      const double x0 = __shfl_sync(full, xb, tau & 31);   // only meaningful while tau < len
      const double up = __shfl_up_sync(full, out, 1);
      const double in = (lane == 0) ? x0 : up;
      const int tloc = tau - lane;
      if (lane < nseg && tloc >= 0 && tloc < len) {
        // ref: musk_model.py:178-187 Calc_Discharge_V1 — (c0*new[i] + c1*old[i]) + c2*old[i+1]
        const double nv = c0 * in + c1 * prev_in + c2 * prev_out;
        prev_in = in;
        prev_out = nv;
        out = nv;
      }
      const double y = __shfl_sync(full, out, nseg - 1);
      const int p = tau - (nseg - 1);
      if (p == lane) yb = y;
    }
    if (lane < nseg) dis_new[g0 + cbase + lane + 1] = prev_out;
    x_own = yb;                                      // the next segment group continues on this output
    if (cbase + 32 >= S && lane < len) rout[ta + lane] = yb;
  }
}

// Wide anti-diagonals (thousands of reaches, e.g. all headwater reaches at once) are throughput-bound, and there the
// systolic warp-per-reach mapping wastes lanes (12 segments on average) and ≈25 instructions per tick.  This kernel
// maps LANE = REACH instead: a warp takes 32 reaches, stages their 32-step inflow tile in shared memory with coalesced
// row loads (lanes over time), then every lane runs its own reach sequentially over time and segments — the same
// operations in the same order ((c0*new[i] + c1*old[i]) + c2*old[i+1], ref: musk_model.py:178-187), ≈0.3 warp
// instructions per reach-segment-step — and the outflow tile goes back with coalesced stores.  The inflow rows were
// written by reach_diag_kernel(bulk = 1) in the launch before.
__global__ void __launch_bounds__(32 * HB_BULK_WARPS) reach_bulk_kernel(const RiverDev d, const int32_t* __restrict__ reaches,
                                                                       int n, int diag, int n_chunks) {
  extern __shared__ double smem[];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* X = smem + (size_t)wib * (32 * 33 + (HB_BULK_SMAX + 1) * 32);   // [reach lane][33]: inflow, then outflow
  double* Q = X + 32 * 33;                                              // [segment point][lane]
  const unsigned full = 0xffffffffu;
  const int task = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32 + lane;
  int64_t r = -1, ta = 0, g0 = 0;
  int len = 0, S = 0;
  double c0 = 0.0, c1 = 0.0, c2 = 0.0;
  const int64_t nt = d.nt;
  if (task < n) {
    const int64_t rr = reaches[task];
    const int c = diag - d.reach_level[rr];
    g0 = d.seg_ptr[rr];
    S = (int)(d.seg_ptr[rr + 1] - g0 - 1);
    if (c >= 0 && c < n_chunks && S >= 1 && S <= HB_BULK_SMAX) {
      r = rr;
      ta = (int64_t)c * HB_RIVER_CHUNK;
      const int64_t tb = (ta + HB_RIVER_CHUNK < nt) ? ta + HB_RIVER_CHUNK : nt;
      len = (int)(tb - ta);
      c0 = d.coef[0 * d.n_reach + r]; c1 = d.coef[1 * d.n_reach + r]; c2 = d.coef[2 * d.n_reach + r];
    }
  }
  if (__all_sync(full, r < 0)) return;
  const int c_own = (r >= 0) ? diag - d.reach_level[r] : 0;
  // ---- inflow tile: row j of the tile = reach of lane j, lanes over time (coalesced) -----------------------------------
#pragma unroll 4
  for (int j = 0; j < 32; ++j) {
    const int64_t rj = __shfl_sync(full, r, j);
    const int64_t taj = __shfl_sync(full, ta, j);
    const int lenj = __shfl_sync(full, len, j);
    if (rj >= 0 && lane < lenj) X[j * 33 + lane] = d.reach_in[rj * nt + taj + lane];
  }
  // ---- state of the own reach ----------------------------------------------------------------------------------------------
  if (r >= 0) {
    const double* dis_old = d.dis[(d.par0 + c_own) & 1];
    for (int i = 0; i <= S; ++i) Q[i * 32 + lane] = dis_old[g0 + i];
  }
  __syncwarp();
  // ---- lane = reach: sequential over time and segments ------------------------------------------------------------------
  if (r >= 0) {
    for (int tt = 0; tt < len; ++tt) {
      double in = X[lane * 33 + tt];
      double q_old = Q[lane];
      Q[lane] = in;
      for (int i = 0; i < S; ++i) {
        const double q_next = Q[(i + 1) * 32 + lane];
        const double nv = c0 * in + c1 * q_old + c2 * q_next;
        Q[(i + 1) * 32 + lane] = nv;
        q_old = q_next;
        in = nv;
      }
      X[lane * 33 + tt] = in;
    }
    double* dis_new = d.dis[(d.par0 + c_own + 1) & 1];
    for (int i = 0; i <= S; ++i) dis_new[g0 + i] = Q[i * 32 + lane];
  }
  __syncwarp();
  // ---- outflow tile back to the rows (coalesced) --------------------------------------------------------------------------
#pragma unroll 4
  for (int j = 0; j < 32; ++j) {
    const int64_t rj = __shfl_sync(full, r, j);
    const int64_t taj = __shfl_sync(full, ta, j);
    const int lenj = __shfl_sync(full, len, j);
    if (rj >= 0 && lane < lenj) d.reach_out[rj * nt + taj + lane] = X[j * 33 + lane];
  }
}

}  // namespace hb
