// river_kernel.cuh — musk_classic routing over the node/reach DAG as a topological-level wavefront.
//
// Replaces `c_musk_classic.pyx` (`Pick_Inflow_V1`, `Update_Discharge_V1`, `Calc_Discharge_V1`, `Calc_Outflow_V1`,
// `Pass_Outflow_V1`; ref: hydpy/models/musk/musk_model.py:19-187, 809-848; SegmentModel.run
// hydpy/core/modeltools.py:3449-3461) and the node summation of the reference's whole-period mode
// (ref: hydpy/core/threadingtools.py:396-446).
//
// All series are device rows [row][nt] (time contiguous):  land rows (outlets.q of the land elements), node rows,
// reach inflow/outflow rows, external rows.  Per level two launches:
//   node_sum_kernel : one warp per node, lanes over time, entries added in the listed order (0.0 + a + b + …);
//   reach_kernel    : one warp per reach; lane i owns segment i and the warp runs as a systolic pipeline over
//                     "ticks": at tick τ lane i computes time τ-i, receiving new[i](t) from lane i-1 by shuffle.
//                     The recurrence  new[i+1] = c0*new[i] + c1*old[i] + c2*old[i+1]  is evaluated with the
//                     reference's operation order and no FMA contraction ⇒ bit-identical to the CPU path.
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
  const double* dis_old;    // [n_seg] discharge at window start
  double* dis_new;          // [n_seg] discharge at window end
};

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

// value a node hands to its downstream reach at time t (ref: threadingtools.py:433-446)
__device__ __forceinline__ double delivered(const RiverDev& d, int32_t node, int64_t t) {
  const int mode = d.node_mode[node];
  double v;
  if (mode == HB_NODE_NEWSIM) v = d.node_rows[(int64_t)node * d.nt + t];
  else {
    v = d.ext_rows[(int64_t)d.node_ext[node] * d.nt + t];
    if (mode == HB_NODE_OBSNEWSIM && isnan(v)) v = d.node_rows[(int64_t)node * d.nt + t];
  }
  return 0.0 + v;
}

__global__ void __launch_bounds__(128) reach_kernel(const RiverDev d, const int32_t* __restrict__ reaches, int n) {
  const int warp = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (warp >= n) return;
  const unsigned full = 0xffffffffu;
  const int64_t r = reaches[warp];
  const int64_t nt = d.nt;
  const int64_t g0 = d.seg_ptr[r];
  const int S = (int)(d.seg_ptr[r + 1] - g0 - 1);
  const double c0 = d.coef[0 * d.n_reach + r], c1 = d.coef[1 * d.n_reach + r], c2 = d.coef[2 * d.n_reach + r];
  const int64_t j0 = d.inlet_ptr[r], j1 = d.inlet_ptr[r + 1];
  double* rin = d.reach_in + r * nt;
  double* rout = d.reach_out + r * nt;

  // ---- inflow row: ref musk_model.py:26-31 Pick_Inflow_V1 (0.0 + q[0] + q[1] + …) ----------------------------------
  for (int64_t t = lane; t < nt; t += 32) {
    double inflow = 0.0;
    for (int64_t j = j0; j < j1; ++j) inflow += delivered(d, d.inlet_node[j], t);
    rin[t] = inflow;
    if (S == 0) rout[t] = inflow;  // ref: musk_model.py:95-98, 831-835 — zero segments: outflow = discharge[0]
  }
  __syncwarp();
  if (S == 0) {
    if (lane == 0) d.dis_new[g0] = (nt > 0) ? rin[nt - 1] : d.dis_old[g0];
    return;
  }
  if (lane == 0) d.dis_new[g0] = (nt > 0) ? rin[nt - 1] : d.dis_old[g0];

  // ---- systolic pipeline over chunks of 32 segments ------------------------------------------------------------------
  for (int cbase = 0; cbase < S; cbase += 32) {
    const int nseg = (S - cbase < 32) ? (S - cbase) : 32;
    const double* X = (cbase == 0) ? rin : rout;   // later chunks continue in place on the outflow row
    double prev_in = 0.0, prev_out = 0.0, out = 0.0;
    if (lane < nseg) {
      prev_in = d.dis_old[g0 + cbase + lane];
      prev_out = d.dis_old[g0 + cbase + lane + 1];
    }
    const int64_t nticks = nt + nseg - 1;
    double yb = 0.0;
    for (int64_t tau0 = 0; tau0 < nticks; tau0 += 32) {
      const double xb = (tau0 + lane < nt) ? X[tau0 + lane] : 0.0;
      const int jn = (nticks - tau0 < 32) ? (int)(nticks - tau0) : 32;
      for (int j = 0; j < jn; ++j) {
        const int64_t tau = tau0 + j;
        const double x0 = __shfl_sync(full, xb, j);
        const double up = __shfl_up_sync(full, out, 1);
        const double in = (lane == 0) ? x0 : up;
        const int64_t tloc = tau - lane;
        if (lane < nseg && tloc >= 0 && tloc < nt) {
          // ref: musk_model.py:178-187 Calc_Discharge_V1 — (c0*new[i] + c1*old[i]) + c2*old[i+1]
          const double nv = c0 * in + c1 * prev_in + c2 * prev_out;
          prev_in = in;
          prev_out = nv;
          out = nv;
        }
        const double y = __shfl_sync(full, out, nseg - 1);
        const int64_t p = tau - (nseg - 1);
        if (p >= 0) {
          if ((p & 31) == lane) yb = y;
          if ((p & 31) == 31 || p == nt - 1) {
            const int64_t pb = p - (p & 31);
            if (lane <= (p & 31)) rout[pb + lane] = yb;
          }
        }
      }
    }
    __syncwarp();
    if (lane < nseg) d.dis_new[g0 + cbase + lane + 1] = prev_out;
  }
}

}  // namespace hb
