// river_kernel.cuh — musk_classic routing over the node/reach DAG.
//
// Replaces `c_musk_classic.pyx` (`Pick_Inflow_V1`, `Update_Discharge_V1`, `Calc_Discharge_V1`, `Calc_Outflow_V1`,
// `Pass_Outflow_V1`; ref: hydpy/models/musk/musk_model.py:19-187, 809-848; SegmentModel.run
// hydpy/core/modeltools.py:3449-3461) and the node summation of the reference's whole-period mode
// (ref: hydpy/core/threadingtools.py:396-446).
//
// All series are device rows [row][nt] (time contiguous): land rows (outlets.q of the land elements), node rows, reach
// inflow/outflow rows, external rows.  The unit of work is a TASK = (reach, chunk of HB_RIVER_CHUNK = 32 time steps):
// it needs the same chunk of the reaches that feed its inlet nodes and the previous chunk of its own reach (the
// Muskingum state), i.e. tasks on anti-diagonal d = level + chunk only depend on diagonals < d.
//
//   reach_chunk()      one warp, one task.  Lanes over time form the sums of the inlet nodes (entries added in the
//                      listed order, 0.0 + a + b + …) and store the rows of the nodes this reach owns; then lane i owns
//                      segment i and the chunk runs as a systolic pipeline over "ticks": at tick τ lane i computes time
//                      τ-i, receiving new[i](t) from lane i-1 by shuffle.  The recurrence
//                      new[i+1] = c0*new[i] + c1*old[i] + c2*old[i+1] is evaluated with the reference's operation order
//                      and no FMA contraction ⇒ bit-identical to the CPU path.
//   reach_head_kernel  level-0 reaches (headwaters: nothing upstream but land elements — half of a river tree): all
//                      their tasks are independent, one launch covers every (reach, chunk); reaches with 1..32 segments
//                      only get their inflow here and
//   reach_bulk_kernel  runs their recurrences with LANE = REACH: the 32 reaches of a warp keep their segment states in
//                      registers, inflow/outflow tiles move through shared memory with coalesced row accesses, ≈0.3 warp
//                      instructions per reach-segment-step instead of ≈25 per tick of the systolic mapping.
//   reach_wave_kernel  everything else in ONE persistent launch: warps draw tasks from an ordered queue (diagonal-major,
//                      a valid topological order), wait on per-reach progress counters for the tasks they depend on and
//                      publish their own.  Because a task only waits for tasks drawn EARLIER — by warps that are resident
//                      and running — the scheme cannot deadlock whatever share of the GPU the kernel gets; a bounded spin
//                      turns any violation of that invariant into an error flag instead of a hang.  Sequential depth:
//                      (levels + chunks) task latencies instead of as many kernel launches.
//   reach_diag_kernel  the same sweep as one launch per anti-diagonal (HB_RIVER_LAUNCHES): reference implementation of the
//                      wavefront for tests and fallback.
//   node_sum_kernel    one warp per node, lanes over time — for the nodes without a downstream reach (outlets).
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
  double* dis[2];           // chunk c reads dis[(par0 + c) & 1] and writes the other one
  int par0;
  int n_chunks;
  const int32_t* reach_level;  // [n_reach]
  const int32_t* node_owner;   // [n_node] reach that stores the node row (first exit), -1: terminal node
  // persistent wavefront
  const int32_t* wave_order;   // reaches of the wavefront sorted by level
  const int32_t* wave_lvl_ptr; // [n_levels+1] into wave_order
  const int32_t* diag_start;   // [n_diag+1] prefix sums of the tasks per anti-diagonal
  const int64_t* dep_ptr;      // [n_reach+1] → dep_reach: the reaches feeding the inlet nodes of a reach
  const int32_t* dep_reach;
  int n_levels, n_diag;
  int* counter;                // task queue head
  int* progress;               // [n_reach] chunks completed in this run
  int* error;                  // set when a wait ran into the spin limit
};

#define HB_RIVER_CHUNK 32
#define HB_BULK_SMAX 32      // reaches with more segments always take the systolic path
#define HB_BULK_WARPS 4      // warps per block of reach_bulk_kernel
#define HB_BULK_SMEM (HB_BULK_WARPS * 32 * 33 * 8)
#define HB_SPIN_LIMIT (1 << 24)

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

// ref: threadingtools.py:396-415 node series = 0.0 + entries in order.  Reach outflow rows may have been written by
// another SM moments ago (persistent wavefront): they are read through L2 (ld.global.cg), never from a stale L1 line.
__device__ __forceinline__ double node_sum_at(const RiverDev& d, int64_t node, int64_t t) {
  double acc = 0.0;
  const int64_t i0 = d.entry_ptr[node], i1 = d.entry_ptr[node + 1];
  for (int64_t i = i0; i < i1; ++i) {
    const int32_t s = d.entry_src[i];
    if (s >= 0) acc += d.land_rows[(int64_t)s * d.nt + t];
    else if (s > HB_ENTRY_EXT) acc += __ldcg(d.reach_out + (int64_t)(-1 - s) * d.nt + t);
    else acc += d.ext_rows[(int64_t)(HB_ENTRY_EXT - s) * d.nt + t];
  }
  return acc;
}

__global__ void __launch_bounds__(256) node_sum_kernel(const RiverDev d, const int32_t* __restrict__ nodes, int n) {
  const int warp = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (warp >= n) return;
  const int64_t node = nodes[warp];
  double* row = d.node_rows + node * d.nt;
  for (int64_t t = lane; t < d.nt; t += 32) row[t] = node_sum_at(d, node, t);
}

// One task by one warp.  bulk: reaches with 1..HB_BULK_SMAX segments stop after the inflow (reach_bulk_kernel continues);
// concurrent: several chunks of this reach run at the same time (head kernel) — only the last one stores the state of
// a segment-free reach, straight into the final buffer.
__device__ __forceinline__ void reach_chunk(const RiverDev& d, int64_t r, int c, int lane, bool bulk, bool concurrent) {
  const unsigned full = 0xffffffffu;
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
  if (bulk && S >= 1 && S <= HB_BULK_SMAX) return;
  const double x_last = __shfl_sync(full, x_own, len - 1);
  if (lane == 0) {
    if (!concurrent) dis_new[g0] = x_last;
    else if (c == d.n_chunks - 1) d.dis[(d.par0 + d.n_chunks) & 1][g0] = x_last;
  }
  if (S == 0) return;

  // ---- systolic pipeline over groups of 32 segments --------------------------------------------------------------------
  for (int cbase = 0; cbase < S; cbase += 32) {
    const int nseg = (S - cbase < 32) ? (S - cbase) : 32;
    double prev_in = 0.0, prev_out = 0.0, out = 0.0;
    if (lane < nseg) {   // through L2: the previous chunk may have been written by another SM (persistent wavefront)
      prev_in = __ldcg(dis_old + g0 + cbase + lane);
      prev_out = __ldcg(dis_old + g0 + cbase + lane + 1);
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

// ---- one launch per anti-diagonal ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) reach_diag_kernel(const RiverDev d, const int32_t* __restrict__ reaches, int n,
                                                         int diag) {
  const int warp = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (warp >= n) return;
  const int64_t r = reaches[warp];
  const int c = diag - d.reach_level[r];
  if (c < 0 || c >= d.n_chunks) return;
  reach_chunk(d, r, c, lane, false, false);
}

// ---- headwater reaches: every (reach, chunk) at once ------------------------------------------------------------------------
__global__ void __launch_bounds__(128) reach_head_kernel(const RiverDev d, const int32_t* __restrict__ reaches, int n) {
  const int64_t task = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (task >= (int64_t)n * d.n_chunks) return;
  const int c = (int)(task / n);
  reach_chunk(d, reaches[task - (int64_t)c * n], c, lane, true, true);
}

// lane = reach; `reaches` = the head reaches with 1..HB_BULK_SMAX segments (their inflow rows are complete)
__global__ void __launch_bounds__(32 * HB_BULK_WARPS) reach_bulk_kernel(const RiverDev d, const int32_t* __restrict__ reaches,
                                                                       int n) {
  extern __shared__ double smem[];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* X = smem + (size_t)wib * (32 * 33);        // [reach lane][33]: inflow, then outflow of the chunk
  const unsigned full = 0xffffffffu;
  const int task = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32 + lane;
  const int64_t nt = d.nt;
  int64_t r = -1, g0 = 0;
  int S = 0;
  double c0 = 0.0, c1 = 0.0, c2 = 0.0;
  double q[HB_BULK_SMAX + 1];                        // discharge at the segment points: registers (static indexing only)
  if (task < n) {
    r = reaches[task];
    g0 = d.seg_ptr[r];
    S = (int)(d.seg_ptr[r + 1] - g0 - 1);
    c0 = d.coef[0 * d.n_reach + r]; c1 = d.coef[1 * d.n_reach + r]; c2 = d.coef[2 * d.n_reach + r];
  }
  const double* dis_old = d.dis[d.par0 & 1];
#pragma unroll
  for (int i = 0; i <= HB_BULK_SMAX; ++i) q[i] = (r >= 0 && i <= S) ? dis_old[g0 + i] : 0.0;
  int smax = S;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const int v = __shfl_xor_sync(full, smax, o); smax = v > smax ? v : smax; }
  for (int c = 0; c < d.n_chunks; ++c) {
    const int64_t ta = (int64_t)c * HB_RIVER_CHUNK;
    const int len = (int)((ta + HB_RIVER_CHUNK < nt ? ta + HB_RIVER_CHUNK : nt) - ta);
    // inflow tile: row j = reach of lane j, lanes over time (coalesced)
#pragma unroll 8
    for (int j = 0; j < 32; ++j) {
      const int64_t rj = __shfl_sync(full, r, j);
      if (rj >= 0 && lane < len) X[j * 33 + lane] = d.reach_in[rj * nt + ta + lane];
    }
    __syncwarp();
    if (r >= 0) {
      for (int tt = 0; tt < len; ++tt) {
        double in = X[lane * 33 + tt];
        double q_old = q[0];
        q[0] = in;
#pragma unroll
        for (int i = 0; i < HB_BULK_SMAX; ++i) {
          if (i >= smax) break;                       // warp-uniform
          if (i < S) {
            // ref: musk_model.py:178-187 Calc_Discharge_V1 — (c0*new[i] + c1*old[i]) + c2*old[i+1]
            const double q_next = q[i + 1];
            const double nv = c0 * in + c1 * q_old + c2 * q_next;
            q[i + 1] = nv;
            q_old = q_next;
            in = nv;
          }
        }
        X[lane * 33 + tt] = in;
      }
    }
    __syncwarp();
#pragma unroll 8
    for (int j = 0; j < 32; ++j) {
      const int64_t rj = __shfl_sync(full, r, j);
      if (rj >= 0 && lane < len) d.reach_out[rj * nt + ta + lane] = X[j * 33 + lane];
    }
    __syncwarp();
  }
  double* dis_fin = d.dis[(d.par0 + d.n_chunks) & 1];
#pragma unroll
  for (int i = 0; i <= HB_BULK_SMAX; ++i)
    if (r >= 0 && i <= S) dis_fin[g0 + i] = q[i];
}

// ---- persistent wavefront ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// queue head and error flag = 0; progress = 0 everywhere, then = n_chunks for the reaches the head kernels complete
__global__ void wave_init_kernel(const RiverDev d) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) { *d.counter = 0; *d.error = 0; }
  if (i < d.n_reach) d.progress[i] = 0;
}
__global__ void wave_head_done_kernel(const RiverDev d, const int32_t* __restrict__ head, int n_head) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_head) d.progress[head[i]] = d.n_chunks;
}

__global__ void __launch_bounds__(128) reach_wave_kernel(const RiverDev d, int n_tasks) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  for (;;) {
    int task = 0;
    if (lane == 0) task = atomicAdd(d.counter, 1);
    task = __shfl_sync(full, task, 0);
    if (task >= n_tasks) return;
    // anti-diagonal of the task: largest dg with diag_start[dg] <= task
    int lo = 0, hi = d.n_diag;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (d.diag_start[mid] <= task) lo = mid; else hi = mid;
    }
    const int dg = lo;
    const int l_lo = dg - d.n_chunks + 1 > 0 ? dg - d.n_chunks + 1 : 0;
    const int64_t r = d.wave_order[d.wave_lvl_ptr[l_lo] + (task - d.diag_start[dg])];
    const int c = dg - d.reach_level[r];
    // ---- wait: same chunk of every feeding reach, previous chunk of this reach (lanes over dependencies) -------------
    bool ok = true;
    {
      const int64_t p0 = d.dep_ptr[r], p1 = d.dep_ptr[r + 1];
      for (int64_t p = p0 + lane - 1; p < p1; p += 32) {     // p0-1 stands for the reach itself
        const int* flag = d.progress + (p < p0 ? r : (int64_t)d.dep_reach[p]);
        const int need = p < p0 ? c : c + 1;
        int spins = 0;
        while (ld_acquire(flag) < need) {
          __nanosleep(100);
          if (++spins > HB_SPIN_LIMIT || ld_acquire(d.error) != 0) { ok = false; break; }
        }
      }
    }
    if (!__all_sync(full, ok)) {
      if (lane == 0) atomicExch(d.error, 1);
      return;
    }
    reach_chunk(d, r, c, lane, false, false);
    __threadfence();
    __syncwarp();
    if (lane == 0) st_release(d.progress + r, c + 1);
  }
}

}  // namespace hb
