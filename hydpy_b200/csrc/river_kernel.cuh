// river_kernel.cuh — musk_classic (and musk_mct, see "musk_mct" below) routing over the node/reach DAG.
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

// Everything a task needs to know about its reach in one 64-byte record (one cache line, one load latency)
struct ReachMeta {
  double c0, c1, c2;     // musk_control.Coefficients
  int32_t g0;            // first point of the reach in the discharge state
  int32_t nseg;          // NmbSegments
  int32_t level;
  int32_t dep0, n_dep;   // → dep_reach: the reaches feeding the inlet nodes
  int32_t simple_node;   // >= 0: the reach has ONE inlet node, in deploy mode newsim, with <= 32 entries (fast inflow)
  int32_t entry0, n_entry;
  int32_t owns_node;     // 1: this reach stores the row of simple_node
  int32_t pad;           // 1: musk_mct reach (parameters in RiverDev::rpar/tpar, conditions in mct_cn/mct_rn/mct_rwd)
};
static_assert(sizeof(ReachMeta) == 64, "ReachMeta must stay one 64-byte record");

struct RiverDev {
  int64_t n_node, n_reach, nt;
  const ReachMeta* meta;    // [n_reach]
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
  double* dis_series;       // optional [nt][n_seg]: musk_classic states.discharge at every step (NULL: not recorded)
  int64_t n_seg;
  double* dis[2];           // chunk c reads dis[(par0 + c) & 1] and writes the other one
  int par0;
  int n_chunks;
  const int32_t* reach_level;  // [n_reach]
  const int32_t* node_owner;   // [n_node] reach that stores the node row (first exit), -1: terminal node
  // persistent wavefront
  const int32_t* wave_rec;     // [n_wave][32] one 128-byte record per wavefront reach, in wave order (see HB_WR_*)
  const int32_t* wave_order;   // reaches of the wavefront sorted by level
  const int32_t* wave_lvl_ptr; // [n_levels+1] into wave_order
  const int32_t* diag_start;   // [n_diag+1] prefix sums of the tasks per anti-diagonal
  const int64_t* dep_ptr;      // [n_reach+1] → dep_reach: the reaches feeding the inlet nodes of a reach
  const int32_t* dep_reach;
  int n_levels, n_diag;
  int* counter;                // task queue head
  int* progress;               // [n_reach] chunks completed in this run, times pscale
  int pscale;                  // 1; grouped wavefront: HB_GRP_NSUB (its groups publish every sub-chunk of HB_GRP_SUB steps)
  int* error;                  // set when a wait ran into the spin limit
  // grouped wavefront (reach_wave2_kernel): units = groups of up to 32 simple reaches of one level, or single reaches
  const int32_t* unit_start;   // [n_units+1] into unit_reach
  const int32_t* unit_reach;   // member reaches, unit after unit
  const int32_t* unit_info;    // [n_units] level | (group ? 1 << 30 : 0)
  const int32_t* unit_lvl_ptr; // [n_levels+1] into the units (sorted by level)
  const int32_t* task_list;    // [n_units * n_chunks] unit | chunk << 24, in wavefront order
  const double* zero_row;      // 32 zeros
  double* mct_series;          // optional [popc(mct_series_mask)][nt][n_seg]: 1-D musk_mct sequences at every step
  unsigned mct_series_mask;    // bits of enum hb_mct_series; slot of a sequence = number of lower bits set
  // musk_mct reaches
  const int32_t* nmbruns;      // [n_reach]
  const double* rpar;          // [HB_RP_COUNT][n_reach]
  const int64_t* trap_ptr;     // [n_reach+1]
  const double* tpar;          // [HB_TP_COUNT][n_trap]
  int64_t n_trap;
  double* mct_cn;              // [n_seg] states.courantnumber (updated in place: one task per reach at a time)
  double* mct_rn;              // [n_seg] states.reynoldsnumber
  double* mct_rwd;             // [n_seg] factors.referencewaterdepth
};

#define HB_RIVER_CHUNK 32
#define HB_BULK_SMAX 32      // reaches with more segments always take the systolic path
#define HB_BULK_WARPS 4      // warps per block of reach_bulk_kernel
#define HB_BULK_SMEM (HB_BULK_WARPS * 32 * 33 * 8)
#define HB_SPIN_LIMIT (1 << 24)
#ifndef HB_SPIN_FAST
#define HB_SPIN_FAST 4
#endif
#ifndef HB_SPIN_SLEEP
#define HB_SPIN_SLEEP 256
#endif
// Polls of tasks that wait long (the narrow tail of a river tree offers fewer tasks than the kernel has warps; ncu, r02a:
// 22 % of the wavefront's instructions are polls of tasks waiting for their own previous chunk) could pause longer:
// HB_SPIN_MAXSHIFT > 0 doubles the pause every HB_SPIN_STEP polls up to HB_SPIN_SLEEP << HB_SPIN_MAXSHIFT ns.  Measured
// (profiles/r02i_spin_backoff.txt, 438-h windows): the runoff generation beside it gains 0.2-0.6 ms, but the wavefront
// itself wakes up late on its critical path (6.2 -> 7.2 / 9.8 ms alone for 4 / 16 us) and the pipelined step LOSES
// (18.7 -> 20.2 / 20.8 ms).  The default stays a constant pause.
#ifndef HB_SPIN_STEP
#define HB_SPIN_STEP 4
#endif
#ifndef HB_SPIN_MAXSHIFT
#define HB_SPIN_MAXSHIFT 0
#endif
__device__ __forceinline__ void spin_pause(int spins) {
  if (spins > HB_SPIN_FAST) {
    int sh = (spins - HB_SPIN_FAST) / HB_SPIN_STEP;
    sh = sh > HB_SPIN_MAXSHIFT ? HB_SPIN_MAXSHIFT : sh;
    __nanosleep((unsigned)HB_SPIN_SLEEP << sh);
  }
}

// qt [nt][e_pad]  →  land rows [n_elem][nt]   (32x32 shared-memory tile transpose, both sides coalesced)
// (elements e_begin .. n_elem-1 with a grid of ceil((n_elem - e_begin) / 32) x ceil(nt / 32) blocks)
__global__ void __launch_bounds__(256) transpose_qt_kernel(const double* __restrict__ qt, double* __restrict__ rows,
                                                           int64_t nt, int64_t n_elem, int64_t e_pad, int64_t e_begin) {
  __shared__ double tile[32][33];
  const int64_t e0 = e_begin + (int64_t)blockIdx.x * 32, t0 = (int64_t)blockIdx.y * 32;
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

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int ld_relaxed(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Wait until the progress counter *flag has reached `need`.  Most of a task's life in the persistent wavefront is this
// loop (ncu, r02a: 83 polls per task, half of the kernel's instructions), and its issue slots are taken from the runoff
// generation of the next window on the same SMs — so the poll is a relaxed load (an `ld.acquire` per poll also invalidates
// the SM's L1: CCTL.IVALL, 13 % of the kernel's stall samples), the bookkeeping of the bounded spin runs every 1024th
// poll, and ONE acquire load of the same counter after the loop orders the reads that follow.  False: spin limit / error.
#ifndef HB_WAVE_POLL
#define HB_WAVE_POLL 2   // 1: round 1's form (an ld.acquire and the full bookkeeping per poll), for A/B timing
#endif
__device__ __forceinline__ bool wait_progress(const int* flag, int need, const int* error) {
#if HB_WAVE_POLL == 2
  if (ld_relaxed(flag) < need) {
    int spins = 0;
    for (;;) {
      if (spins >= HB_SPIN_FAST) __nanosleep(HB_SPIN_SLEEP);
      if (ld_relaxed(flag) >= need) break;
      if ((++spins & 1023) == 0 && (spins > HB_SPIN_LIMIT || ld_relaxed(error) != 0)) return false;
    }
  }
  (void)ld_acquire(flag);
  return true;
#else
  int spins = 0;
  while (ld_acquire(flag) < need) {
    spin_pause(++spins);
    if (spins > HB_SPIN_LIMIT || ((spins & 1023) == 0 && ld_acquire(error) != 0)) return false;
  }
  return true;
#endif
}

// ---- musk_mct (ref: hydpy/models/musk_mct.py RUN_METHODS; cross section hydpy/models/wq_trapeze_strickler.py) ----------
// The Muskingum-Cunge-Todini segment update has the data flow of the classic one (new[i], old[i], old[i+1]), so it runs
// in the same systolic pipeline: lane i = segment i keeps its Courant/Reynolds numbers and the reference water depth in
// registers; the coefficients follow from a Pegasus root search for the reference water depth per segment and step.
#define HB_PYMAX_R(a, b) (((b) > (a)) ? (b) : (a))
#define HB_PYMIN_R(a, b) (((b) < (a)) ? (b) : (a))
struct WqDev {
  int n;
  const double *wb, *ss, *ks, *cf, *bd, *ht, *ws, *pd;
  double bottomslope;
};
struct WqRes { double wettedarea, surfacewidth, discharge, celerity; };

// ref: wq_model.py Use_WaterDepth_V1 :2176-2252 = Calc_WettedAreas_V1 :580-597, Calc_WettedPerimeters_V1 :936-954,
// Calc_WettedPerimeterDerivatives_V1 :1178-1189, Calc_SurfaceWidths_V1 :1371-1383, Calc_Discharges_V1 :1632-1649,
// Calc_DischargeDerivatives_V1 :1860-1882, Calc_Celerity_V1 :2069-2075 and the sums over the trapezes
__device__ __noinline__ void wq_use_waterdepth(const WqDev& q, double waterdepth, WqRes& o) {
  double wettedarea = 0.0, surfacewidth = 0.0, discharge = 0.0, dischargederivative = 0.0;
  const double sqrt_bs = pow(q.bottomslope, 0.5);
  for (int i = 0; i < q.n; ++i) {
    const double wb = __ldg(q.wb + i), ss = __ldg(q.ss + i), ht = __ldg(q.ht + i), ws = __ldg(q.ws + i), bd = __ldg(q.bd + i);
    const double d = waterdepth - bd;
    double a, p, dp, w;
    if (d < 0.0) { a = 0.0; p = 0.0; dp = 0.0; w = 0.0; }
    else if (d < ht) {
      a = (wb + ss * d) * d;
      p = wb + 2.0 * d * pow(ss * ss + 1.0, 0.5);   // (reference: pow(pow(ss, 2.0) + 1.0, 0.5); the square is exact either way)
      dp = __ldg(q.pd + i);
      w = wb + 2.0 * ss * d;
    } else {
      a = (wb + ws / 2.0) * ht + (wb + ws) * (d - ht);
      p = wb + 2.0 * ht * pow(ss * ss + 1.0, 0.5) + 2.0 * (d - ht);
      dp = 2.0;
      w = wb + ws;
    }
    wettedarea += a;
    surfacewidth += w;
    if (waterdepth > bd) {
      const double k = __ldg(q.cf + i) * __ldg(q.ks + i) * sqrt_bs;
      discharge += k * pow(a, 5.0 / 3.0) / pow(p, 2.0 / 3.0);
      dischargederivative += k * pow(a / p, 2.0 / 3.0) * (5.0 * p * w - 2.0 * a * dp) / (3.0 * p);
    } else { discharge += 0.0; dischargederivative += 0.0; }
  }
  o.wettedarea = wettedarea;
  o.surfacewidth = surfacewidth;
  o.discharge = discharge;
  o.celerity = (surfacewidth > 0.0) ? dischargederivative / surfacewidth : __longlong_as_double(0x7ff8000000000000LL);
}

// ref: cythons/rootutils.pyx:25-83 PegasusBase.find_x on Return_ReferenceDischargeError_V1 (musk_model.py:306-314)
__device__ __noinline__ double mct_find_waterdepth(const WqDev& q, double refq, double x0, double x1, double xmin,
                                                   double xmax, double xtol, double ytol, int itermax) {
  WqRes tmp;
  double x = __longlong_as_double(0x7ff8000000000000LL), y, y0, y1, dx, y0_abs, y1_abs;
  if (x0 > x1) { const double t = x0; x0 = x1; x1 = t; }
  if (xmin > xmax) { const double t = xmin; xmin = xmax; xmax = t; }
  x0 = HB_PYMAX_R(x0, xmin);
  x1 = HB_PYMIN_R(x1, xmax);
  for (int guard = 0; guard < 4096; ++guard) {
    wq_use_waterdepth(q, x0, tmp); y0 = tmp.discharge - refq;
    y0_abs = fabs(y0);
    if (y0_abs <= ytol) return x0;
    wq_use_waterdepth(q, x1, tmp); y1 = tmp.discharge - refq;
    y1_abs = fabs(y1);
    if (y1_abs <= ytol) return x1;
    dx = x1 - x0;
    if ((y0 < 0 && y1 < 0) || (y0 > 0 && y1 > 0)) {
      if ((x0 == xmin) && (x1 == xmax)) return (y0_abs <= y1_abs) ? x0 : x1;
      { const double a = x0 - dx; x0 = HB_PYMAX_R(a, xmin); }
      { const double a = x1 + dx; x1 = HB_PYMIN_R(a, xmax); }
    } else {
      guard = -1;
      break;
    }
  }
  if (fabs(dx) < xtol) return (x0 + x1) / 2;
  // a NaN reference discharge leaves the bracket loop at once (no comparison holds) and turns every iterate below into
  // NaN: same result as the reference's itermax rounds (ref: rootutils.pyx:57-78), without evaluating them
  if (!(y0 == y0) || !(y1 == y1)) return x;
  for (int iter = 0; iter < itermax; ++iter) {
    x = x0 - y0 * dx / (y1 - y0);
    if (x < xmin) return xmin;
    else if (x > xmax) return xmax;
    wq_use_waterdepth(q, x, tmp); y = tmp.discharge - refq;
    if (fabs(y) < ytol) return x;
    if (((y1 < 0) && (y < 0)) || ((y1 > 0) && (y > 0))) y0 *= y1 / (y1 + y);
    else { x0 = x1; y0 = y1; }
    x1 = x;
    y1 = y;
    dx = x1 - x0;
    if (fabs(dx) < xtol) break;
  }
  return x;
}

// One segment, one step (all NmbRuns repetitions): returns new.discharge[i+1]; cn/rn/rwd: old values in, new values out.
// ref: musk_model.py:245-254 Calc_ReferenceDischarge_V1, 383-399 Calc_ReferenceWaterDepth_V1, 415-424, 506-512
// Calc_CorrectingFactor_V1, 548-560 Calc_CourantNumber_V1, 600-616 Calc_ReynoldsNumber_V1, 695-705
// Calc_Coefficient1_Coefficient2_Coefficient3_V1, 792-806 Calc_Discharge_V2
// rec: cell [t][segment] of the first recorded series (NULL: no series requested)
__device__ __noinline__ double mct_segment(const RiverDev& d, int64_t r, double in, double prev_in, double prev_out,
                                           double* st /* cn, rn, rwd */, double* rec) {
  const int64_t nr = d.n_reach, t0 = d.trap_ptr[r];
  WqDev q;
  q.n = (int)(d.trap_ptr[r + 1] - t0);
  q.wb = d.tpar + (int64_t)HB_TP_BOTTOMWIDTH * d.n_trap + t0;
  q.ss = d.tpar + (int64_t)HB_TP_SIDESLOPE * d.n_trap + t0;
  q.ks = d.tpar + (int64_t)HB_TP_STRICKLERCOEFFICIENT * d.n_trap + t0;
  q.cf = d.tpar + (int64_t)HB_TP_CALIBRATIONFACTOR * d.n_trap + t0;
  q.bd = d.tpar + (int64_t)HB_TP_BOTTOMDEPTH * d.n_trap + t0;
  q.ht = d.tpar + (int64_t)HB_TP_TRAPEZEHEIGHT * d.n_trap + t0;
  q.ws = d.tpar + (int64_t)HB_TP_SLOPEWIDTH * d.n_trap + t0;
  q.pd = d.tpar + (int64_t)HB_TP_PERIMETERDERIVATIVE * d.n_trap + t0;
  q.bottomslope = d.rpar[(int64_t)HB_RP_WQ_BOTTOMSLOPE * nr + r];
  const double tolwd = d.rpar[(int64_t)HB_RP_TOLERANCEWATERDEPTH * nr + r], tolq = d.rpar[(int64_t)HB_RP_TOLERANCEDISCHARGE * nr + r],
               bottomslope = d.rpar[(int64_t)HB_RP_BOTTOMSLOPE * nr + r], seconds = d.rpar[(int64_t)HB_RP_SECONDS * nr + r],
               segmentlength = d.rpar[(int64_t)HB_RP_SEGMENTLENGTH * nr + r];
  const int nruns = d.nmbruns[r];
  const double cn_old = st[0], rn_old = st[1];
  double cn = cn_old, rn = rn_old, rwd = st[2], nv = prev_out;
  for (int run = 0; run < nruns; ++run) {
    const double est = (run == 0) ? prev_out + in - prev_in : nv;
    const double a = (in + est) / 2.0;
    const double refq = HB_PYMAX_R(a, 0.0);
    double mn, mx;
    if (isnan(rwd) || isinf(rwd)) { mn = 0.0; mx = 2.0; }
    else if (rwd <= 0.001) { mn = 0.0; mx = 0.01; }
    else { mn = 0.9 * rwd; mx = 1.1 * rwd; }
    const double b = refq / 10.0;
    const double tol_q = HB_PYMIN_R(tolq, b);
    rwd = mct_find_waterdepth(q, refq, mn, mx, 0.0, 1000.0, tolwd, tol_q, 100);
    WqRes w;
    wq_use_waterdepth(q, rwd, w);
    double corr;
    if (refq == 0.0) { corr = 1.0; cn = 0.0; rn = 0.0; }
    else {
      corr = w.celerity * w.wettedarea / refq;
      cn = (w.celerity / corr) * (seconds / (1000.0 * segmentlength));
      rn = refq / (corr * w.surfacewidth * bottomslope * w.celerity * (1000.0 * segmentlength));
    }
    double f = 1.0 / (1.0 + cn + rn);
    const double c1 = (cn + rn - 1.0) * f;
    if (cn_old != 0.0) f *= cn / cn_old;
    const double c2 = (1 + cn_old - rn_old) * f;
    const double c3 = (1 - cn_old + rn_old) * f;
    if (in == prev_in && prev_in == prev_out) nv = in;
    else nv = c1 * in + c2 * prev_in + c3 * prev_out;
    nv = HB_PYMAX_R(nv, 0.0);
    if (rec && run == nruns - 1) {
      // what save_data finds after the last repetition (ref: modelutils.py:1127-1167)
      const double vals[HB_MS_COUNT] = {refq, rwd, w.wettedarea, w.surfacewidth, w.celerity, corr, c1, c2, c3, cn, rn};
      const int64_t stride = d.nt * d.n_seg;
      int slot = 0;
#pragma unroll
      for (int k = 0; k < HB_MS_COUNT; ++k)
        if (d.mct_series_mask >> k & 1u) rec[(slot++) * stride] = vals[k];
    }
  }
  st[0] = cn; st[1] = rn; st[2] = rwd;
  return nv;
}

// One task by one warp.  bulk: reaches with 1..HB_BULK_SMAX segments stop after the inflow (reach_bulk_kernel continues);
// concurrent: several chunks of this reach run at the same time (head kernel) — only the last one stores the state of
// a segment-free reach, straight into the final buffer.  WAVE: wait for the tasks this one depends on (persistent
// wavefront) — AFTER all metadata is in registers, so that only the row loads and the ticks sit on the critical path.
// Returns false when the wait ran into the spin limit.
// words of a wavefront record (int32 each; doubles occupy two words)
enum { HB_WR_C0 = 0, HB_WR_C1 = 2, HB_WR_C2 = 4, HB_WR_REACH = 6, HB_WR_G0, HB_WR_NSEG, HB_WR_LEVEL, HB_WR_NDEP,
       HB_WR_NENTRY, HB_WR_NODE, HB_WR_OWNS, HB_WR_INLINE, HB_WR_PAD, HB_WR_ENTRY = 16, HB_WR_DEP = 24, HB_WR_WORDS = 32 };
#define HB_WR_MAXLIST 8
#define HB_NO_ENTRY 0x7fffffff

// what a task knows about its reach before it touches any series
struct TaskPrep {
  ReachMeta m;
  int64_t r;
  int32_t esrc;    // entry_src of lane e of the simple inlet node (HB_NO_ENTRY otherwise)
  int32_t dep;     // reach lane e waits for (wavefront, inline list), -1: none
  bool inline_deps;
};

__device__ __forceinline__ TaskPrep prep_from_meta(const RiverDev& d, int64_t r, int lane) {
  TaskPrep p;
  p.m = d.meta[r];
  p.r = r;
  p.esrc = (p.m.simple_node >= 0 && lane < p.m.n_entry) ? d.entry_src[p.m.entry0 + lane] : HB_NO_ENTRY;
  p.dep = -1;
  p.inline_deps = false;
  return p;
}

// MCT: the instance also handles musk_mct reaches (only launched when the network holds some: the calls into the root
// search cost the lean classic instance its register budget — 56 registers, which is what lets the persistent
// wavefront share the SMs with the runoff generation of the next window)
template <bool WAVE, bool MCT>
__device__ __forceinline__ bool reach_chunk(const RiverDev& d, const TaskPrep& tp, int c, int lane, bool bulk,
                                            bool concurrent) {
  const unsigned full = 0xffffffffu;
  const int64_t nt = d.nt;
  const ReachMeta& m = tp.m;
  const int64_t r = tp.r;
  const int64_t ta = (int64_t)c * HB_RIVER_CHUNK;
  const int64_t tb = (ta + HB_RIVER_CHUNK < nt) ? ta + HB_RIVER_CHUNK : nt;
  const int len = (int)(tb - ta);
  const double* dis_old = d.dis[(d.par0 + c) & 1];
  double* dis_new = d.dis[(d.par0 + c + 1) & 1];
  const int64_t g0 = m.g0;
  const int S = m.nseg;
  const double c0 = m.c0, c1 = m.c1, c2 = m.c2;
  double* rin = d.reach_in + r * nt;
  double* rout = d.reach_out + r * nt;
  // fast inflow: lane e holds the row of entry e of the single inlet node
  const double* erow = nullptr;
  bool ereach = false;
  if (tp.esrc != HB_NO_ENTRY) {
    const int32_t s = tp.esrc;
    ereach = s < 0 && s > HB_ENTRY_EXT;
    erow = (s >= 0 ? d.land_rows + (int64_t)s * nt
                   : ereach ? d.reach_out + (int64_t)(-1 - s) * nt : d.ext_rows + (int64_t)(HB_ENTRY_EXT - s) * nt) + ta;
  }
  // state of the first segment group; in the wavefront it is loaded as soon as the previous chunk of THIS reach is
  // done — that happened a whole diagonal ago — so that the load overlaps the wait for the upstream reaches
  double st_in = 0.0, st_out = 0.0;
  if (WAVE) {
    bool ok = true;
    if (lane == 0) ok = wait_progress(d.progress + r, c * d.pscale, d.error);
    if (!__all_sync(full, ok)) return false;
    __syncwarp();     // acquire side: the polling lane's ld.acquire is ordered before every lane's loads below
    if (lane < S) {   // through L2: the previous chunk may have been written by another SM
      st_in = __ldcg(dis_old + g0 + lane);
      st_out = __ldcg(dis_old + g0 + lane + 1);
    }
    // same chunk of every feeding reach (lanes over dependencies)
    for (int p = lane; p < m.n_dep; p += 32)
      ok = ok && wait_progress(d.progress + (tp.inline_deps ? tp.dep : d.dep_reach[m.dep0 + p]), (c + 1) * d.pscale, d.error);
    if (!__all_sync(full, ok)) return false;
    __syncwarp();
  } else if (lane < S) {
    st_in = dis_old[g0 + lane];
    st_out = dis_old[g0 + lane + 1];
  }

  // ---- inflow of the chunk (lane = time step): node sums, node rows of owned nodes, Pick_Inflow_V1 ----------------------
  double x_own = 0.0;   // inflow at time ta+lane, kept in the register for the pipeline below
  if (m.simple_node >= 0) {
    double sim = 0.0;   // ref: threadingtools.py:396-415 node series = 0.0 + entries in order
    for (int e = 0; e < m.n_entry; ++e) {
      const double* row = (const double*)__shfl_sync(full, (unsigned long long)erow, e);
      const bool isr = __shfl_sync(full, (int)ereach, e) != 0;
      if (lane < len) sim += isr ? __ldcg(row + lane) : row[lane];
    }
    if (lane < len) {
      const int64_t t = ta + lane;
      if (m.owns_node) d.node_rows[(int64_t)m.simple_node * nt + t] = sim;
      const double inflow = 0.0 + (0.0 + sim);   // ref: musk_model.py:26-31 Pick_Inflow_V1
      rin[t] = inflow;
      if (S == 0) rout[t] = inflow;
      if (d.dis_series) d.dis_series[t * d.n_seg + g0] = inflow;
      x_own = inflow;
    }
  } else if (lane < len) {
    const int64_t j0 = d.inlet_ptr[r], j1 = d.inlet_ptr[r + 1];
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
    if (d.dis_series) d.dis_series[t * d.n_seg + g0] = inflow;
    x_own = inflow;
  }
  const bool mct = MCT && m.pad != 0;
  if (mct && lane < len) {
    // ref: musk_model.py:67-74 Adjust_Inflow_V1 (fluxes.inflow holds the adjusted value)
    if (x_own < 0.0) {
      const double tolneg = d.rpar[(int64_t)HB_RP_TOLERANCENEGATIVEINFLOW * d.n_reach + r];
      x_own = (x_own < tolneg) ? __longlong_as_double(0x7ff8000000000000LL) : 0.0;
      const int64_t t = ta + lane;
      rin[t] = x_own;
      if (S == 0) rout[t] = x_own;
      if (d.dis_series) d.dis_series[t * d.n_seg + g0] = x_own;
    }
  }
  if (bulk && S >= 1 && S <= HB_BULK_SMAX) return true;
  const double x_last = __shfl_sync(full, x_own, len - 1);
  if (lane == 0) {
    if (!concurrent) dis_new[g0] = x_last;
    else if (c == d.n_chunks - 1) d.dis[(d.par0 + d.n_chunks) & 1][g0] = x_last;
  }
  if (S == 0) return true;

  // ---- systolic pipeline over groups of 32 segments --------------------------------------------------------------------
  for (int cbase = 0; cbase < S; cbase += 32) {
    const int nseg = (S - cbase < 32) ? (S - cbase) : 32;
    double prev_in = st_in, prev_out = st_out, out = 0.0;
    if (cbase > 0) {
      prev_in = prev_out = 0.0;
      if (lane < nseg) {
        prev_in = __ldcg(dis_old + g0 + cbase + lane);
        prev_out = __ldcg(dis_old + g0 + cbase + lane + 1);
      }
    }
    double mst[3] = {0.0, 0.0, 0.0};                 // musk_mct: Courant number, Reynolds number, reference water depth
    if (mct && lane < nseg) {
      mst[0] = __ldcg(d.mct_cn + g0 + cbase + lane);
      mst[1] = __ldcg(d.mct_rn + g0 + cbase + lane);
      mst[2] = __ldcg(d.mct_rwd + g0 + cbase + lane);
    }
    double xb = x_own;                               // input of this segment group at time ta+lane
    double yb = 0.0;                                 // its output at time ta+lane (collected below)
    const int nticks = len + nseg - 1;
    // (≈40 warp instructions per tick for five flops: everything that does not depend on the tick stays out of the loop)
    const bool record = d.dis_series != nullptr;
// (HB_WAVE_TICK = 2, measured and rejected: the last segment's lane stores the outflow row tick by tick instead of the
// shuffle that collects it into lane = time step — 6.2 -> 9.0 ms for the whole network)
#ifndef HB_WAVE_TICK
#define HB_WAVE_TICK 1
#endif
    const bool single = HB_WAVE_TICK == 2 && S <= 32;   // one segment group: its last lane stores the outflow row itself
    const bool tail = lane == nseg - 1;
    for (int tau = 0; tau < nticks; ++tau) {
      const double x0 = __shfl_sync(full, xb, tau & 31);   // only meaningful while tau < len
      const double up = __shfl_up_sync(full, out, 1);
      const double in = (lane == 0) ? x0 : up;
      const int tloc = tau - lane;
      if (lane < nseg && tloc >= 0 && tloc < len) {
        // ref: musk_model.py:178-187 Calc_Discharge_V1 — (c0*new[i] + c1*old[i]) + c2*old[i+1]
        const double nv = mct ? mct_segment(d, r, in, prev_in, prev_out, mst,
                                            d.mct_series ? d.mct_series + (ta + tloc) * d.n_seg + g0 + cbase + lane : nullptr)
                              : c0 * in + c1 * prev_in + c2 * prev_out;
        prev_in = in;
        prev_out = nv;
        out = nv;
        if (record) d.dis_series[(ta + tloc) * d.n_seg + g0 + cbase + lane + 1] = nv;
        if (single && tail) rout[ta + tloc] = nv;
      }
      if (!single) {
        const double y = __shfl_sync(full, out, nseg - 1);
        const int p = tau - (nseg - 1);
        if (p == lane) yb = y;
      }
    }
    if (lane < nseg) dis_new[g0 + cbase + lane + 1] = prev_out;
    if (mct && lane < nseg) {
      d.mct_cn[g0 + cbase + lane] = mst[0];
      d.mct_rn[g0 + cbase + lane] = mst[1];
      d.mct_rwd[g0 + cbase + lane] = mst[2];
    }
    x_own = yb;                                      // the next segment group continues on this output
    if (!single && cbase + 32 >= S && lane < len) rout[ta + lane] = yb;
  }
  return true;
}

// ---- one launch per anti-diagonal ----------------------------------------------------------------------------------------
template <bool MCT>
__global__ void __launch_bounds__(128) reach_diag_kernel(const RiverDev d, const int32_t* __restrict__ reaches, int n,
                                                         int diag) {
  const int warp = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (warp >= n) return;
  const int64_t r = reaches[warp];
  const int c = diag - d.reach_level[r];
  if (c < 0 || c >= d.n_chunks) return;
  reach_chunk<false, MCT>(d, prep_from_meta(d, r, lane), c, lane, false, false);
}

// ---- headwater reaches: every (reach, chunk) at once ------------------------------------------------------------------------
__global__ void __launch_bounds__(128) reach_head_kernel(const RiverDev d, const int32_t* __restrict__ reaches, int n) {
  const int64_t task = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (task >= (int64_t)n * d.n_chunks) return;
  const int c = (int)(task / n);
  reach_chunk<false, false>(d, prep_from_meta(d, reaches[task - (int64_t)c * n], lane), c, lane, true, true);
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
    const ReachMeta m = d.meta[r];
    g0 = m.g0; S = m.nseg; c0 = m.c0; c1 = m.c1; c2 = m.c2;
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
            if (d.dis_series) d.dis_series[(ta + tt) * d.n_seg + g0 + i + 1] = nv;
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
// queue head = 0; progress = 0 everywhere, then = n_chunks for the reaches the head kernels complete
__global__ void wave_init_kernel(const RiverDev d) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *d.counter = 0;   // (the error flag is sticky: hb_wait reports and clears it)
  if (i < d.n_reach) d.progress[i] = 0;
}
__global__ void wave_head_done_kernel(const RiverDev d, const int32_t* __restrict__ head, int n_head) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_head) d.progress[head[i]] = d.n_chunks * d.pscale;
}

// dynamic shared memory: copies of diag_start [n_diag+1] and wave_lvl_ptr [n_levels+1] when they fit (use_smem).
// Per-task latency is a chain of L2 round trips (≈0.7 µs each on B200): queue draw → record → progress flags → rows.
// The first two are taken off the critical path by working two tasks ahead: the draw for task i+2 and the record load
// for task i+1 are in flight while task i runs.  A warp therefore owns up to three tasks, always processed in
// increasing order, which keeps the no-deadlock argument intact (the smallest unfinished task is always some warp's
// current task or becomes it as soon as that warp's smaller tasks — already finished — are retired).
// 9 blocks of 128 threads per SM cap the kernel at 56 registers (what it needs without spills): every register the
// wavefront holds is taken from the runoff generation of the next window (at 76 registers the pipelined step lost 0.3 ms;
// 48/40/32 registers make the routing itself slower than that gains)
#ifndef HB_WAVE_MIN_BLOCKS
#define HB_WAVE_MIN_BLOCKS 9
#endif
template <bool MCT>
__global__ void __launch_bounds__(128, MCT ? 1 : HB_WAVE_MIN_BLOCKS) reach_wave_kernel(const RiverDev d, int n_tasks, int use_smem) {
  extern __shared__ int32_t wave_tab[];
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int32_t* diag_start = d.diag_start;
  const int32_t* lvl_ptr = d.wave_lvl_ptr;
  if (use_smem) {
    for (int i = threadIdx.x; i <= d.n_diag; i += blockDim.x) wave_tab[i] = d.diag_start[i];
    for (int i = threadIdx.x; i <= d.n_levels; i += blockDim.x) wave_tab[d.n_diag + 1 + i] = d.wave_lvl_ptr[i];
    __syncthreads();
    diag_start = wave_tab;
    lvl_ptr = wave_tab + d.n_diag + 1;
  }
  // task index → (anti-diagonal, position in wave order)
  auto locate = [&](int task, int& dg, int& pos) {
    int lo = 0, hi = d.n_diag;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (diag_start[mid] <= task) lo = mid; else hi = mid;
    }
    dg = lo;
    const int l_lo = dg - d.n_chunks + 1 > 0 ? dg - d.n_chunks + 1 : 0;
    pos = lvl_ptr[l_lo] + (task - diag_start[dg]);
  };
  auto load_record = [&](int task, int& dg) -> int {   // one coalesced 128-byte load: lane i holds word i
    if (task >= n_tasks) { dg = 0; return 0; }
    int pos;
    locate(task, dg, pos);
    return d.wave_rec[(int64_t)pos * HB_WR_WORDS + lane];
  };
  auto draw = [&]() -> int {
    int t = 0;
    if (lane == 0) t = atomicAdd(d.counter, 1);
    return t;                                           // valid in lane 0; broadcast when it is needed
  };
  int t_cur = __shfl_sync(full, draw(), 0);
  int t_nxt = __shfl_sync(full, draw(), 0);
  int dg_cur, dg_nxt;
  int w_cur = load_record(t_cur, dg_cur);
  int w_nxt = load_record(t_nxt, dg_nxt);
  while (t_cur < n_tasks) {
    const int a_nn = draw();                            // in flight while the current task runs
    // ---- unpack the record of the current task -----------------------------------------------------------------------
    TaskPrep tp;
#define HB_W(i) __shfl_sync(full, w_cur, (i))
    tp.m.c0 = __hiloint2double(HB_W(HB_WR_C0 + 1), HB_W(HB_WR_C0));
    tp.m.c1 = __hiloint2double(HB_W(HB_WR_C1 + 1), HB_W(HB_WR_C1));
    tp.m.c2 = __hiloint2double(HB_W(HB_WR_C2 + 1), HB_W(HB_WR_C2));
    tp.r = HB_W(HB_WR_REACH);
    tp.m.g0 = HB_W(HB_WR_G0);
    tp.m.nseg = HB_W(HB_WR_NSEG);
    tp.m.level = HB_W(HB_WR_LEVEL);
    const int inline_ok = HB_W(HB_WR_INLINE);
    const int e_word = __shfl_sync(full, w_cur, HB_WR_ENTRY + (lane & (HB_WR_MAXLIST - 1)));
    const int d_word = __shfl_sync(full, w_cur, HB_WR_DEP + (lane & (HB_WR_MAXLIST - 1)));
    if (inline_ok) {
      tp.m.n_dep = HB_W(HB_WR_NDEP);
      tp.m.n_entry = HB_W(HB_WR_NENTRY);
      tp.m.simple_node = HB_W(HB_WR_NODE);
      tp.m.owns_node = HB_W(HB_WR_OWNS);
      tp.m.dep0 = 0; tp.m.entry0 = 0; tp.m.pad = 0;
      tp.esrc = lane < tp.m.n_entry ? e_word : HB_NO_ENTRY;
      tp.dep = lane < tp.m.n_dep ? d_word : -1;
      tp.inline_deps = true;
    } else {
      tp = prep_from_meta(d, tp.r, lane);               // long lists / several inlet nodes / deploy modes: global arrays
    }
#undef HB_W
    const int c = dg_cur - tp.m.level;
    if (!reach_chunk<true, MCT>(d, tp, c, lane, false, false)) {
      if (lane == 0) atomicExch(d.error, 1);
      return;
    }
    // bar.warp.sync orders the lanes' stores before lane 0's release, which is cumulative at gpu scope
    __syncwarp();
    if (lane == 0) st_release(d.progress + tp.r, c + 1);
    // ---- rotate the pipeline ------------------------------------------------------------------------------------------------
    const int t_nn = __shfl_sync(full, a_nn, 0);
    int dg_nn;
    const int w_nn = load_record(t_nn, dg_nn);
    t_cur = t_nxt; w_cur = w_nxt; dg_cur = dg_nxt;
    t_nxt = t_nn; w_nxt = w_nn; dg_nxt = dg_nn;
  }
}

// ---- grouped wavefront ---------------------------------------------------------------------------------------------------
// The warp-per-task wavefront above spends ≈2000 warp instructions on one (reach, chunk) — 32 + S systolic ticks for S
// segments — and needs 4 blocks of 128 threads per SM to keep enough tasks in flight; those blocks take registers and
// issue slots from the runoff generation of the next window, which shares the GPU with the routing (ncu/bench: the
// zone kernel runs 4.0 ms alone, 5.7 ms beside the wavefront).  Here a task is (GROUP OF 32 REACHES of one level, chunk):
// lanes over time build the 32 inflow rows into a shared-memory tile (coalesced row reads, like reach_chunk), then
// lane = reach runs the Muskingum recurrences of its reach with the segment states in registers (like
// reach_bulk_kernel), ≈140 warp instructions per (reach, chunk).  One small block per SM is enough.  Reaches the group
// form does not cover (several inlet nodes, deploy modes, more than HB_GRP_SMAX segments, musk_mct) stay single-reach
// units handled by reach_chunk.  Dependencies, ordering and therefore deadlock freedom are those of reach_wave_kernel:
// tasks are drawn in anti-diagonal order and only wait for tasks drawn earlier.
#define HB_GRP_SMAX 24
#define HB_GRP_EMAX 4       // entries of the inlet node of a group reach (land element + upstream reaches)
#define HB_GRP_SUB 8        // steps of a sub-chunk
#define HB_GRP_NSUB (HB_RIVER_CHUNK / HB_GRP_SUB)
#define HB_GRP_WARPS 2
#ifndef HB_GRP_MIN_BLOCKS
#define HB_GRP_MIN_BLOCKS 6   // caps the kernel at 168 registers
#endif
#define HB_GRP_SMEM (HB_GRP_WARPS * 32 * (HB_GRP_SUB + 1) * 8)

#ifdef HB_GRP_DEBUG
__device__ unsigned long long hb_grp_dbg[8];
#define HB_DBG_T(k) do { if (lane == 0) { const long long now_ = clock64(); atomicAdd(&hb_grp_dbg[k], (unsigned long long)(now_ - t_dbg)); t_dbg = now_; } } while (0)
#else
#define HB_DBG_T(k) do { } while (0)
#endif
__device__ __forceinline__ bool poll_flags(const RiverDev& d, const int* const* flag, int n_slots, int need, bool active) {
  // converged polling: one warp-wide relaxed load per flag slot and round, ONE acquire fence at the end
  const unsigned full = 0xffffffffu;
  int spins = 0;
  for (;;) {
    bool ready = true;
    if (active)
#pragma unroll
      for (int p = 0; p < HB_GRP_EMAX; ++p)
        if (p < n_slots && flag[p] != nullptr) ready = ready && (ld_relaxed(flag[p]) >= need);
    if (__all_sync(full, ready)) break;
    if (++spins > HB_SPIN_FAST) __nanosleep(HB_SPIN_SLEEP);
    if (spins > HB_SPIN_LIMIT || ((spins & 1023) == 0 && ld_relaxed(d.error) != 0)) return false;   // warp-uniform
  }
  __threadfence();   // acquire side of the flags' release stores
  return true;
}

// One (group, 32-step chunk) task.  The chunk runs as HB_GRP_NSUB sub-chunks of HB_GRP_SUB steps, each of which waits only
// for the same sub-chunk of the feeding reaches and publishes itself: consecutive levels of the network overlap inside a
// chunk, so the latency per level is that of a sub-chunk.  progress[r] counts sub-chunks (d.pscale = HB_GRP_NSUB).
__device__ __forceinline__ bool reach_group_chunk(const RiverDev& d, int64_t r, int c, int lane, double* __restrict__ X) {
  const unsigned full = 0xffffffffu;
  const int64_t nt = d.nt;
  const int64_t ta = (int64_t)c * HB_RIVER_CHUNK;
  const int len = (int)((ta + HB_RIVER_CHUNK < nt ? ta + HB_RIVER_CHUNK : nt) - ta);
  const double* dis_old = d.dis[(d.par0 + c) & 1];
  double* dis_new = d.dis[(d.par0 + c + 1) & 1];
  ReachMeta m;
  m.c0 = m.c1 = m.c2 = 0.0; m.g0 = 0; m.nseg = 0; m.level = 0; m.dep0 = 0; m.n_dep = 0; m.simple_node = -1; m.entry0 = 0;
  m.n_entry = 0; m.owns_node = 0; m.pad = 0;
#ifdef HB_GRP_DEBUG
  long long t_dbg = clock64();
#endif
  if (r >= 0) m = d.meta[r];
  const int S = m.nseg;
  const int64_t g0 = m.g0;
  HB_DBG_T(0);
  // flags: the own reach (previous chunk) and the feeding reaches; entry rows (slots beyond the entry list point to a
  // row of zeros, so that every load below is unconditional)
  const int* own_flag[HB_GRP_EMAX] = {r >= 0 ? d.progress + r : nullptr, nullptr, nullptr, nullptr};
  const int* dep_flag[HB_GRP_EMAX];
  const double* erow[HB_GRP_EMAX];
#pragma unroll
  for (int p = 0; p < HB_GRP_EMAX; ++p) {
    dep_flag[p] = (r >= 0 && p < m.n_dep) ? d.progress + __ldg(d.dep_reach + m.dep0 + p) : nullptr;
    erow[p] = d.zero_row;
    if (r >= 0 && p < m.n_entry) {
      const int32_t s = __ldg(d.entry_src + m.entry0 + p);
      erow[p] = (s >= 0 ? d.land_rows + (int64_t)s * nt
                        : s > HB_ENTRY_EXT ? d.reach_out + (int64_t)(-1 - s) * nt : d.ext_rows + (int64_t)(HB_ENTRY_EXT - s) * nt) + ta;
    }
  }
  if (!poll_flags(d, own_flag, 1, c * HB_GRP_NSUB, r >= 0)) return false;
  HB_DBG_T(1);
  double q[HB_GRP_SMAX + 1];                          // discharge at the segment points (static indexing only)
#pragma unroll
  for (int i = 0; i <= HB_GRP_SMAX; ++i) q[i] = (r >= 0 && i <= S) ? __ldcg(dis_old + g0 + i) : 0.0;
  int smax = S;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const int v = __shfl_xor_sync(full, smax, o); smax = v > smax ? v : smax; }
  const double c0 = m.c0, c1 = m.c1, c2 = m.c2;
  const int jr = lane >> 3, ts = lane & (HB_GRP_SUB - 1);   // inflow/row phases: lane = (reach within a batch of 4, step)
  HB_DBG_T(2);
  for (int sub = 0; sub < HB_GRP_NSUB; ++sub) {
    const int slen = len - sub * HB_GRP_SUB < HB_GRP_SUB ? len - sub * HB_GRP_SUB : HB_GRP_SUB;
    if (slen <= 0) break;                               // warp-uniform
    const int64_t t0 = ta + sub * HB_GRP_SUB;
    const bool last = (sub + 1) * HB_GRP_SUB >= len;
    if (!poll_flags(d, dep_flag, HB_GRP_EMAX, c * HB_GRP_NSUB + sub + 1, r >= 0)) return false;
    HB_DBG_T(1);
    // ---- inflow tile: 8 batches of 4 reaches x 8 steps, all 32 row loads in flight ------------------------------------
    // (ref: threadingtools.py:396-415 node series = 0.0 + entries in order; musk_model.py:26-31 Pick_Inflow_V1)
    const int ts_c = ts < slen ? ts : slen - 1;
    double v[8][HB_GRP_EMAX];
#pragma unroll
    for (int jb = 0; jb < 8; ++jb)
#pragma unroll
      for (int e = 0; e < HB_GRP_EMAX; ++e) {
        const double* row = (const double*)__shfl_sync(full, (unsigned long long)erow[e], jb * 4 + jr);
        v[jb][e] = __ldcg(row + sub * HB_GRP_SUB + ts_c);   // through L2: rows of other SMs' reaches
      }
#pragma unroll
    for (int jb = 0; jb < 8; ++jb) {
      const int j = jb * 4 + jr;
      const int64_t rj = __shfl_sync(full, r, j);
      const int nodej = __shfl_sync(full, m.simple_node, j), ownsj = __shfl_sync(full, m.owns_node, j);
      const int Sj = __shfl_sync(full, S, j);
      const int64_t g0j = __shfl_sync(full, g0, j);
      // slots beyond the entry list hold +0.0, and x + 0.0 == x for every x this sum can take (it starts at +0.0)
      double sim = 0.0;
#pragma unroll
      for (int e = 0; e < HB_GRP_EMAX; ++e) sim += v[jb][e];
      if (rj >= 0 && ts < slen) {
        const int64_t t = t0 + ts;
        if (ownsj) d.node_rows[(int64_t)nodej * nt + t] = sim;
        const double inflow = 0.0 + (0.0 + sim);
        d.reach_in[rj * nt + t] = inflow;
        if (Sj == 0) d.reach_out[rj * nt + t] = inflow;
        if (d.dis_series) d.dis_series[t * d.n_seg + g0j] = inflow;
        X[j * (HB_GRP_SUB + 1) + ts] = inflow;
      }
    }
    __syncwarp();
    HB_DBG_T(3);
    // ---- lane = reach: the recurrences of the sub-chunk (ref: musk_model.py:178-187 Calc_Discharge_V1) -----------------
    if (r >= 0) {
      for (int tt = 0; tt < slen; ++tt) {
        double in = X[lane * (HB_GRP_SUB + 1) + tt];
        double q_old = q[0];
        q[0] = in;
#pragma unroll
        for (int i = 0; i < HB_GRP_SMAX; ++i) {
          if (i >= smax) break;                         // warp-uniform
          if (i < S) {
            const double q_next = q[i + 1];
            const double nv = c0 * in + c1 * q_old + c2 * q_next;   // (c0*new[i] + c1*old[i]) + c2*old[i+1]
            q[i + 1] = nv;
            q_old = q_next;
            in = nv;
            if (d.dis_series) d.dis_series[(t0 + tt) * d.n_seg + g0 + i + 1] = nv;
          }
        }
        X[lane * (HB_GRP_SUB + 1) + tt] = in;
      }
    }
    __syncwarp();
    HB_DBG_T(4);
#pragma unroll
    for (int jb = 0; jb < 8; ++jb) {
      const int j = jb * 4 + jr;
      const int64_t rj = __shfl_sync(full, r, j);
      const int Sj = __shfl_sync(full, S, j);
      if (rj >= 0 && Sj > 0 && ts < slen) d.reach_out[rj * nt + t0 + ts] = X[j * (HB_GRP_SUB + 1) + ts];
    }
    if (last) {
#pragma unroll
      for (int i = 0; i <= HB_GRP_SMAX; ++i)
        if (r >= 0 && i <= S) dis_new[g0 + i] = q[i];
    }
    // bar.warp.sync orders every lane's stores (rows are written lanes-over-time) before the releases below
    __syncwarp();
    HB_DBG_T(5);
    if (r >= 0) st_release(d.progress + r, last ? (c + 1) * HB_GRP_NSUB : c * HB_GRP_NSUB + sub + 1);
    HB_DBG_T(6);
  }
#ifdef HB_GRP_DEBUG
  if (lane == 0) atomicAdd(&hb_grp_dbg[7], 1ull);
#endif
  return true;
}

template <bool MCT>
__global__ void __launch_bounds__(32 * HB_GRP_WARPS, HB_GRP_MIN_BLOCKS) reach_wave2_kernel(const RiverDev d, int n_tasks) {
  extern __shared__ double grp_smem[];
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  double* X = grp_smem + (size_t)wib * (32 * (HB_GRP_SUB + 1));
  auto draw = [&]() -> int {
    int t = 0;
    if (lane == 0) t = atomicAdd(d.counter, 1);
    return __shfl_sync(full, t, 0);
  };
  int t_cur = draw();
  while (t_cur < n_tasks) {
    const int t_nxt = draw();                           // in flight while the current task runs
    // task → (unit, chunk): the host lists the tasks in the order of the sub-chunk wavefront, level + HB_GRP_NSUB * chunk
    // (a task waits for level - 1 of the same chunk and for the previous chunk of its own level: both earlier in this
    // order), so that the tasks the warps hold are the ones that can run
    const int packed = __ldg(d.task_list + t_cur);
    const int unit = packed & 0xffffff, c = packed >> 24;
    const int info = __ldg(d.unit_info + unit);
    const int us = __ldg(d.unit_start + unit), cnt = __ldg(d.unit_start + unit + 1) - us;
    bool ok;
    if (info >> 30) {
      const int64_t r = lane < cnt ? (int64_t)__ldg(d.unit_reach + us + lane) : -1;
      ok = reach_group_chunk(d, r, c, lane, X);
    } else {
      const int64_t r = __ldg(d.unit_reach + us);
      ok = reach_chunk<true, MCT>(d, prep_from_meta(d, r, lane), c, lane, false, false);
      __syncwarp();
      if (ok && lane == 0) st_release(d.progress + r, (c + 1) * d.pscale);
    }
    if (!ok) {
      if (lane == 0) atomicExch(d.error, 1);
      return;
    }
    t_cur = t_nxt;
  }
}

// ---- goodness-of-fit statistics of node series against observations --------------------------------------------------------
// The sufficient statistics of hydpy/auxs/statstools.py (nse :574-616, kge :944-987, rmse :529-553, corr, bias, std
// ratio) for many (node, observation row) pairs at once — the objective reduction of a calibration ensemble, so that only
// HB_ST_COUNT doubles per pair travel to the host instead of the series.  One block per pair, two passes (means first,
// then centred sums: no cancellation), fixed reduction tree ⇒ deterministic.  skip_nan drops the steps where either value
// is NaN (`prepare_arrays(skip_nan=True)`, statstools.py:480-483); otherwise NaNs propagate like in NumPy.
__device__ __forceinline__ double block_sum(double v, double* red) {
  const unsigned full = 0xffffffffu;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(full, v, o);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
  __syncthreads();
  if (threadIdx.x == 0) red[0] = t;
  __syncthreads();
  return red[0];
}

__global__ void __launch_bounds__(256) node_statistics_kernel(const double* __restrict__ node_rows, const double* __restrict__ obs,
                                                              const int32_t* __restrict__ nodes, const int32_t* __restrict__ obs_row,
                                                              int64_t nt, int skip_nan, double* __restrict__ out) {
  __shared__ double red[8];
  const int pair = blockIdx.x;
  const double* s = node_rows + (int64_t)nodes[pair] * nt;
  const double* o = obs + (int64_t)obs_row[pair] * nt;
  double n = 0.0, ss = 0.0, so = 0.0;
  for (int64_t t = threadIdx.x; t < nt; t += blockDim.x) {
    const double a = s[t], b = o[t];
    if (skip_nan && (isnan(a) || isnan(b))) continue;
    n += 1.0; ss += a; so += b;
  }
  n = block_sum(n, red); ss = block_sum(ss, red); so = block_sum(so, red);
  const double ms = ss / n, mo = so / n;
  double e2 = 0.0, vs = 0.0, vo = 0.0, sp = 0.0;
  for (int64_t t = threadIdx.x; t < nt; t += blockDim.x) {
    const double a = s[t], b = o[t];
    if (skip_nan && (isnan(a) || isnan(b))) continue;
    const double da = a - ms, db = b - mo, d = a - b;
    e2 += d * d; vs += da * da; vo += db * db; sp += da * db;
  }
  e2 = block_sum(e2, red); vs = block_sum(vs, red); vo = block_sum(vo, red); sp = block_sum(sp, red);
  if (threadIdx.x == 0) {
    double* r = out + (int64_t)pair * HB_ST_COUNT;
    r[HB_ST_N] = n; r[HB_ST_MEAN_SIM] = ms; r[HB_ST_MEAN_OBS] = mo; r[HB_ST_SS_ERR] = e2; r[HB_ST_SS_SIM] = vs;
    r[HB_ST_SS_OBS] = vo; r[HB_ST_SP] = sp;
  }
}

}  // namespace hb
