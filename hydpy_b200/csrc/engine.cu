// engine.cu — host side of libhydpy_b200.so: the C-ABI of include/hydpy_b200.h over the sm_100a kernels.
//
// Replaces, for hland_96(+evap_aet_hbv96/evap_pet_hbv96/rconc_uh) and musk_classic, what the reference does in
// `HydPy.simulate` → `simulate_multithreaded` (ref: hydpy/core/hydpytools.py:2769-3123) and
// `Model.simulate_period` (ref: hydpy/cythons/modelutils.py:1625-1634).  There is deliberately NO CPU code path for
// the model equations in this library: without a CUDA device `hb_create` fails.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "land_kernel.cuh"
#include "land_kernel_v2.cuh"
#include "land_kernel_coupled.cuh"
#include "river_kernel.cuh"

using namespace hb;

#ifndef HB_COUPLED_B448
#define HB_COUPLED_B448 2   // blocks per SM the fused capillary-flux kernel is compiled for at <= 13 warps per block
#endif
#ifndef HB_WAVE_BLOCKS_PER_SM
#define HB_WAVE_BLOCKS_PER_SM 3   // persistent wavefront blocks per SM: leaves registers for the land kernels of the next window
#endif
#define HB_MAX_SLABS 8
#define HB_FUSE_ZMAX 12      // fused reduction: one warp per zone of a 32-element group; beyond 12 zones only one block
                             // (16 warps) would fit an SM and the term buffers win (measured on the Lahn ensemble)
// dynamic shared memory of zone_kernel<…, ZW>: constants [14][ZW*32] double2 + term buffers [G][2][ZW][32] doubles
#define HB_FUSE_SMEM(zw) ((size_t)(zw) * 32 * (14 * 16 + ZoneGeom<zw>::G * 2 * 8))
// zone/snow series the fast path of hland_96 stages in its own layout (first block + the evap submodels' block)
#define HB_S_STAGED(s) ((s) < HB_S_FIRST_ELEM || ((s) >= HB_S_FIRST_ZONE2 && (s) < HB_S_FIRST_ELEM2) || \
                       ((s) >= HB_S_FIRST_ZONE3 && (s) < HB_S_SC))
enum { HB_K_BEGIN = -1, HB_K_ZONE = 0, HB_K_ELEMENT = 1, HB_K_GENERAL = 2, HB_K_END = 3, HB_K_OTHER = 4 };

namespace {

thread_local std::string g_create_error;

template <typename T>
struct DBuf {  // device buffer that only grows
  T* p = nullptr;
  size_t cap = 0;
  ~DBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  cudaError_t reserve(size_t n) {
    if (n == 0) n = 1;  // never hand out NULL: empty arrays stay addressable
    if (n <= cap) return cudaSuccess;
    release();
    cudaError_t rc = cudaMalloc((void**)&p, n * sizeof(T));
    if (rc == cudaSuccess) cap = n;
    else p = nullptr;
    return rc;
  }
};

// ---- minimal NCCL binding, resolved at run time (torch ships libnccl.so.2; the engine itself has no link-time
// dependency on it, so single-GPU use never needs NCCL) ---------------------------------------------------------------
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool load(std::string& err) {
    if (handle) return true;
    const char* names[] = {getenv("HYDPY_B200_NCCL"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n || !*n) continue;
      handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (handle) break;
    }
    if (!handle) { err = std::string("cannot load NCCL: ") + dlerror(); return false; }
#define HB_SYM(field, name) \
  *(void**)(&field) = dlsym(handle, name); \
  if (!field) { err = std::string("NCCL symbol missing: ") + name; return false; }
    HB_SYM(GetUniqueId, "ncclGetUniqueId") HB_SYM(CommInitRank, "ncclCommInitRank")
    HB_SYM(CommDestroy, "ncclCommDestroy") HB_SYM(Send, "ncclSend") HB_SYM(Recv, "ncclRecv")
    HB_SYM(AllReduce, "ncclAllReduce") HB_SYM(GroupStart, "ncclGroupStart") HB_SYM(GroupEnd, "ncclGroupEnd")
    HB_SYM(GetErrorString, "ncclGetErrorString")
#undef HB_SYM
    return true;
  }
};
const int kNcclFloat64 = 8, kNcclSum = 0;

}  // namespace


// deep copies of the last basin descriptions (hb_set_members builds the ensemble from them)
struct LandBase {
  bool valid = false;
  hb_land_desc d{};
  std::vector<int64_t> zone_ptr, snow_ptr, sfdist_ptr, uh_ptr, sred_ptr;
  std::vector<int32_t> recstep, eflags, forcing_col, outlet_node, zonetype, zflags, zorder, sred_from, sred_to, model, rconc,
      nash_steps;
  std::vector<double> epar, zpar, sfdist, uh, sred_adjust;
  template <typename T>
  static const T* cp(std::vector<T>& v, const T* src, size_t n) {
    if (!src) { v.clear(); return nullptr; }
    v.assign(src, src + n);
    if (v.empty()) v.resize(1);
    return v.data();
  }
  void keep(const hb_land_desc* s) {
    d = *s;
    const size_t ne = (size_t)s->n_elem;
    d.zone_ptr = cp(zone_ptr, s->zone_ptr, ne + 1); d.snow_ptr = cp(snow_ptr, s->snow_ptr, ne + 1);
    d.sfdist_ptr = cp(sfdist_ptr, s->sfdist_ptr, ne + 1); d.uh_ptr = cp(uh_ptr, s->uh_ptr, ne + 1);
    d.sred_ptr = cp(sred_ptr, s->sred_ptr, ne + 1); d.recstep = cp(recstep, s->recstep, ne);
    d.eflags = cp(eflags, s->eflags, ne); d.forcing_col = cp(forcing_col, s->forcing_col, ne);
    d.outlet_node = cp(outlet_node, s->outlet_node, ne); d.epar = cp(epar, s->epar, (size_t)HB_EP_COUNT * ne);
    d.zonetype = cp(zonetype, s->zonetype, (size_t)s->n_zone); d.zflags = cp(zflags, s->zflags, (size_t)s->n_zone);
    d.zorder = cp(zorder, s->zorder, (size_t)s->n_zone); d.zpar = cp(zpar, s->zpar, (size_t)HB_ZP_COUNT * s->n_zone);
    d.sfdist = cp(sfdist, s->sfdist, (size_t)s->n_sfdist); d.uh = cp(uh, s->uh, (size_t)s->n_uh);
    d.sred_from = cp(sred_from, s->sred_from, (size_t)s->n_sred); d.sred_to = cp(sred_to, s->sred_to, (size_t)s->n_sred);
    d.sred_adjust = cp(sred_adjust, s->sred_adjust, (size_t)s->n_sred); d.model = cp(model, s->model, ne);
    d.rconc = cp(rconc, s->rconc, ne); d.nash_steps = cp(nash_steps, s->nash_steps, ne);
    valid = true;
  }
};
struct RiverBase {
  bool valid = false;
  hb_river_desc d{};
  std::vector<int64_t> entry_ptr, inlet_ptr, seg_ptr, trap_ptr;
  std::vector<int32_t> entry_src, node_mode, node_ext, inlet_node, reach_outlet, reach_model, nmbruns;
  std::vector<double> coef, rpar, tpar;
  void keep(const hb_river_desc* s) {
    d = *s;
    const size_t nn = (size_t)s->n_node, nr = (size_t)s->n_reach;
    d.entry_ptr = LandBase::cp(entry_ptr, s->entry_ptr, nn + 1); d.entry_src = LandBase::cp(entry_src, s->entry_src, (size_t)s->n_entry);
    d.node_mode = LandBase::cp(node_mode, s->node_mode, nn); d.node_ext = LandBase::cp(node_ext, s->node_ext, nn);
    d.inlet_ptr = LandBase::cp(inlet_ptr, s->inlet_ptr, nr + 1); d.inlet_node = LandBase::cp(inlet_node, s->inlet_node, (size_t)s->n_inlet);
    d.reach_outlet = LandBase::cp(reach_outlet, s->reach_outlet, nr); d.seg_ptr = LandBase::cp(seg_ptr, s->seg_ptr, nr + 1);
    d.coef = LandBase::cp(coef, s->coef, 3 * nr); d.reach_model = LandBase::cp(reach_model, s->reach_model, nr);
    d.nmbruns = LandBase::cp(nmbruns, s->nmbruns, nr); d.rpar = LandBase::cp(rpar, s->rpar, (size_t)HB_RP_COUNT * nr);
    d.trap_ptr = LandBase::cp(trap_ptr, s->trap_ptr, s->trap_ptr ? nr + 1 : 0);
    d.tpar = LandBase::cp(tpar, s->tpar, (size_t)HB_TP_COUNT * (size_t)s->n_trap);
    valid = true;
  }
};

struct hb_engine {
  int device = 0;
  // ensemble members (hb_set_members): the land and river descriptions on the device are `members` copies of the base
  int64_t members = 1, river_members = 1, member_node_stride = 0;
  LandBase base_land;
  RiverBase base_river;
  // Three streams form the window pipeline: runoff generation of window w+1 (stream) overlaps routing of window w
  // (s_river) and the device→host copy of its node series (s_copy).  All per-window buffers exist twice; `cur` is the
  // set of the window staged last (it flips in hb_set_forcing / hb_select_window).
  cudaStream_t stream = nullptr, s_river = nullptr, s_copy = nullptr, s_comm = nullptr, s_h2d = nullptr;
  // forcing is double-buffered as well: the block of the next window is copied (s_h2d) while this window's land part runs
  int fcur = 0;
  cudaEvent_t ev_forcing_ready[2] = {nullptr, nullptr};   // H2D copy into forcing buffer f is complete
  cudaEvent_t ev_forcing_read[2] = {nullptr, nullptr};    // the last land run that reads forcing buffer f is complete
  // pinned staging of the seasonal factor: a ring, so that the host may enqueue several windows ahead of the device (with
  // one slot per forcing buffer it waited for the copy of two windows ago at every window — a caller whose thread is
  // delayed now and then by the OS could not build up a queue to bridge it)
  static const int SIN_RING = 8;
  double* h_sin[SIN_RING] = {};
  size_t h_sin_cap_ring[SIN_RING] = {};
  cudaEvent_t ev_sin_copied[SIN_RING] = {};
  unsigned sin_next = 0;
  cudaEvent_t ev_ext_ready[2] = {nullptr, nullptr};   // boundary rows of the window have arrived (NCCL, s_comm)
  cudaEvent_t ev_send_done[2] = {nullptr, nullptr};   // boundary rows of the window have left (their buffers may be reused)
  bool ext_async[2] = {false, false};                 // … and the routing of that window has to wait for them
  bool ext_feeds_reach = true;                        // false: external rows only enter nodes without a downstream reach
  int cur = 0;
  cudaEvent_t ev_land_done[2] = {nullptr, nullptr}, ev_river_done[2] = {nullptr, nullptr},
              ev_copy_done[2] = {nullptr, nullptr}, ev_join[2] = {nullptr, nullptr};
  bool pending = false;               // work enqueued by hb_run_async that hb_wait has not yet collected
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_stop = nullptr;
  // per-kernel timing: marks recorded between launches, resolved by hb_wait and accumulated in `timing` until
  // hb_get_timing reads (and resets) the sums
  struct Marks {
    std::vector<cudaEvent_t> ev;
    std::vector<int> tag;             // kernel family that ran BEFORE the mark (HB_K_*), HB_K_BEGIN for an opening mark
    size_t n = 0;
    int add(int t, cudaStream_t st) {
      if (n == ev.size()) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return -1;
        ev.push_back(e);
        tag.push_back(t);
      }
      tag[n] = t;
      return cudaEventRecord(ev[n++], st) == cudaSuccess ? 0 : -1;
    }
  } land_marks, river_marks, comm_marks;   // (comm_marks: only with HB_DEBUG_TIMELINE)
  int mark(int tag) { return land_marks.add(tag, stream); }
  std::string error;
  cudaDeviceProp prop{};

  // ---- land ----
  bool have_land = false, have_states = false;
  int64_t n_elem = 0, e_pad = 0, n_zone = 0, n_snow = 0, n_uh = 0;
  int zmax = 0, cmax = 0, umax = 0;
  bool any_snowlimit = false;
  bool any_nonsoil_warp = false;      // some warp of the zone kernel holds a fast-path zone that is not FIELD/FOREST with FC > 0
  int64_t n_fast = 0;                 // elements eligible for the fast path (HB_EI_FAST)
  int64_t n_fast_classes = 0;         // … of them with several snow classes (fast only while no series is recorded)
  // speculative fast path for finite SMax (HB_EI_SPEC, HB_OPT_LAND_SPECULATE): the elements, the flags their zone kernel
  // raises when a snow class exceeds its limit, and the conditions of the window's start for the repetition
  int land_speculate = 1;
  std::vector<int32_t> spec_list;
  DBuf<int32_t> d_spec_list, d_spec_flag;
  DBuf<double> d_snap_ic, d_snap_sm, d_snap_sp, d_snap_wc, d_snap_uz, d_snap_lz, d_snap_quh;
  std::vector<int32_t> coupled_list;  // elements of the fused kernel (HB_EI_COUPLED)
  std::vector<int32_t> general_list;  // elements of the general kernel in the current combination of paths
  // sibling models: hland_96p / hland_96c elements always run through the general kernel, instantiated
  // for their model
  std::vector<int32_t> h_model, all_siblings[2], sibling_list[2];   // all / those the general kernel takes in this run
  DBuf<int32_t> d_sibling_list[2], d_nash_steps;
  DBuf<double> d_zx1, d_zx2, d_ex1, d_ex2;   // hland_96p suz, sg1, sg2, sg3 | hland_96c bw1, bw2
  bool any_sibling = false, any_nash = false;
  int64_t n_fast_sib[2] = {0, 0};            // their elements eligible for the two-kernel fast path
  bool any_nonsoil_sib[2] = {false, false};
  std::vector<int32_t> h_einfo;
  int general_key = -1;
  DBuf<int32_t> d_coupled_list, d_general_list;
  DBuf<double> d_term_a, d_term_b;    // zone terms of one time chunk, [chunk][zmax][E_pad]
  DBuf<double> d_red_a, d_red_b;      // fused reduction: two sums per element-step of one time chunk, [chunk][E_pad]
  DBuf<int32_t> d_gwarps;             // [E_pad/32] zone warps of a fused 32-element group, 0 = group not fused
  int64_t n_fused = 0;                // elements with HB_EI_FUSED
  int land_fuse = 1;                  // HB_OPT_LAND_FUSE: add the zone terms inside the zone kernel where a group allows
  DBuf<unsigned long long> d_counters;   // HB_CNT_*: what the fast-path kernels executed since the last hb_get_timing
  DBuf<double> d_series_stage[HB_S_COUNT];   // chunk-local staging of the zone/snow series (fast path)
  DBuf<int32_t> d_zone_slot, d_snow_slot;         // ABI cell → k*E_pad + e (fast-path elements), -1 otherwise
  DBuf<int64_t> d_zone_cell0, d_snow_cell0;       // [E_pad+1] first ABI cell of every element
  int land_path = 0;                  // HB_LAND_AUTO / HB_LAND_GENERAL
  // element slabs of the fast path: slab 0 runs on `stream`, slab s on s_slab[s-1]; s_join collects them
  int land_slabs = 2;
  // steps per time chunk of the fast path (0 = as long as the scratch memory allows).  Kernels of ≈1-3 ms instead of
  // ≈5-10 ms: the thread-per-element kernel holds its share of every SM's registers until its LAST block retires, and a
  // kernel that needs a whole CTA slot (the NCCL send/receive of a boundary hand-over) waits that long — measured on 2 GPUs
  // (profiles/r02h_timeline.txt): the boundary rows arrived 7.4 ms late, the downstream rank lost 2.2 ms per step
  // (default -1: 128 steps once the engine takes part in a multi-GPU run, else as long as the memory allows: one GPU
  // alone measured 18.7-18.9 ms per 438-h step unchunked against 18.9-19.8 ms in chunks of 110)
  int64_t land_chunk = -1;
  cudaStream_t s_slab[HB_MAX_SLABS - 1] = {}, s_join = nullptr;
  // the thread-per-element kernel of a slab runs on a stream of higher priority than the thread-per-zone kernels: when a
  // slab's zone kernel ends, its element kernel and the next slab's zone kernel become ready together, and the blocks of
  // the (latency-bound, few-warps) element kernel must reach the SMs first — the zone blocks then fill what is left.  In
  // the other order two zone blocks take an SM's shared memory and the element kernel waits for the whole zone kernel
  // (measured, two slabs, 438 h: 14.8 instead of 11.0 ms, at random from process to process)
  cudaStream_t s_elem[HB_MAX_SLABS][4] = {};   // [slab][fused hland_96 | hland_96 | hland_96p | hland_96c]: these kernels follow
                                               // sequential chains — side by side, never one after the other
  cudaEvent_t ev_slab_zdone[HB_MAX_SLABS] = {}, ev_slab_edone[HB_MAX_SLABS][4] = {};
  cudaEvent_t ev_slab_fork = nullptr, ev_slab_zone[HB_MAX_SLABS] = {}, ev_slab_done[HB_MAX_SLABS] = {};
  int slabs_used = 1;                 // slabs of the current run_land
  int64_t slab_elem_lo[HB_MAX_SLABS + 1] = {};
  int river_sweep = HB_RIVER_PERSISTENT;  // HB_RIVER_PERSISTENT / HB_RIVER_LAUNCHES
  bool river_state_series = false;        // record musk_classic states.discharge per step
  int wave_blocks = HB_WAVE_BLOCKS_PER_SM;  // persistent wavefront blocks per SM
  bool wave_blocks_set = false;             // (the grouped wavefront defaults to 2 blocks of 64 threads)
  DBuf<double> d_dis_series[2];
  unsigned mct_series_mask = 0;           // musk_mct sequences to record per step (bits of enum hb_mct_series)
  unsigned mct_series_run[2] = {0, 0};    // mask the window in buffer set 0/1 was routed with
  DBuf<double> d_mct_series[2];           // [popc(mask)][nt][n_seg]
  std::vector<int64_t> zone_ptr, snow_ptr, uh_ptr;
  std::vector<int32_t> h_nz, h_nc, h_nu, outlet_node;
  DBuf<int32_t> d_nz, d_nc, d_nu, d_recstep, d_einfo, d_fcol, d_zinfo, d_zorder, d_sred_from, d_sred_to;
  DBuf<int64_t> d_zone0, d_snow0, d_uh0, d_sred_ptr;
  DBuf<double> d_kz, d_ke, d_sfdist, d_uh, d_sred_adjust;
  DBuf<double> d_ic, d_sm, d_sp, d_wc, d_uz, d_lz, d_quh, d_scr_a, d_scr_b, d_scr_sred;
  bool series_on[HB_S_COUNT] = {};
  DBuf<double> d_series[HB_S_COUNT];

  // ---- window ----
  int64_t nt = 0, n_col = 0, nt_total = 0, t0 = 0;
  bool have_forcing = false, have_results = false, have_routed = false;
  DBuf<double> d_forcing[2], d_sinfactor[2], d_qt, d_land_rows[2];
  DBuf<MathTables> d_tables;
  DBuf<ChainTables> d_chain_tables;

  // ---- river ----
  bool have_river = false;
  int64_t n_node = 0, n_reach = 0, n_seg = 0, n_ext = 0, n_levels = 0;
  std::vector<int32_t> level_node_ptr, level_reach_ptr;  // [n_levels+1]
  DBuf<int64_t> d_entry_ptr, d_inlet_ptr, d_seg_ptr;
  DBuf<int32_t> d_entry_src, d_node_mode, d_node_ext, d_inlet_node, d_node_order, d_reach_order, d_sel;
  DBuf<int32_t> d_reach_level, d_node_owner, d_terminal_nodes;
  int64_t n_terminal = 0;
  // persistent wavefront + headwater kernels (river_kernel.cuh)
  DBuf<int32_t> d_head_all, d_head_bulk, d_wave_order, d_wave_lvl_ptr, d_dep_reach, d_diag_start;
  DBuf<int64_t> d_dep_ptr;
  DBuf<ReachMeta> d_reach_meta;
  DBuf<int32_t> d_wave_rec;
  DBuf<double> d_obs, d_stats;
  DBuf<int32_t> d_stat_nodes, d_stat_rows;
  DBuf<int> d_progress, d_wave_ctl;     // d_wave_ctl = {queue head, error flag}
  std::vector<int32_t> wave_lvl_ptr;    // [n_levels+1]
  int64_t n_head_all = 0, n_head_bulk = 0, n_wave = 0;
  std::vector<int32_t> plain_ptr_all, plain_ptr_bulk;   // [n_plain_levels+1] into head_all / head_bulk
  int n_plain_levels = 0;
  int64_t river_plain_min = 4096;       // HB_OPT_RIVER_PLAIN_MIN (read by hb_set_river)
  // grouped wavefront (HB_RIVER_GROUPED)
  DBuf<int32_t> d_unit_start, d_unit_reach, d_unit_info, d_unit_lvl_ptr, d_diag_start_g;
  DBuf<double> d_zero_row;
  std::vector<int32_t> unit_lvl_ptr;
  int64_t n_units = 0;
  int diag_chunks_g = -1;
  int group_min_width = 256;            // levels with fewer wavefront reaches run as single-reach (systolic) units
  int diag_chunks = -1;                 // n_chunks the uploaded diag_start belongs to
  int* h_wave_err = nullptr;            // pinned copy of the wavefront's error flag
  // musk_mct reaches
  bool any_mct = false;
  DBuf<int32_t> d_nmbruns;
  DBuf<int64_t> d_trap_ptr;
  DBuf<double> d_rpar, d_tpar, d_mct_cn, d_mct_rn, d_mct_rwd;
  int64_t n_trap = 0;
  DBuf<double> d_coef, d_dis[2], d_ext_rows[2], d_node_rows[2], d_reach_in[2], d_reach_out[2], d_gather[2];
  int dis_cur = 0;
  bool have_ext = false;

  // ---- comm ----
  NcclApi nccl;
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  DBuf<double> d_commbuf;
  // boundary rows travel packed: one message per peer and window.  Staging buffers per (window buffer set, call of the
  // window); `comm_calls` counts the send/receive calls since the window was staged
  std::vector<DBuf<double>> d_comm_stage[2];
  std::vector<DBuf<int32_t>> d_comm_rows[2];
  std::vector<std::vector<int32_t>> comm_rows_host[2];   // what d_comm_rows holds (the same lists come back every window)
  int comm_calls[2] = {0, 0};

  hb_timing timing{};

  int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    error = buf;
    return code;
  }
};

#define HB_CUDA(h, call)                                                                              \
  do {                                                                                                \
    cudaError_t rc__ = (call);                                                                        \
    if (rc__ != cudaSuccess)                                                                          \
      return (h)->fail(rc__ == cudaErrorMemoryAllocation ? HB_ENOMEM : HB_ECUDA, "%s failed: %s (%s:%d)", #call, \
                       cudaGetErrorString(rc__), __FILE__, __LINE__);                                 \
  } while (0)

#define HB_REQUIRE(h, cond, code, ...) \
  do { if (!(cond)) return (h)->fail(code, __VA_ARGS__); } while (0)


// ---- folding of the reference's parameters into the constants the kernels work with ----------------------------------
// Evaluated exactly as the reference evaluates the sub-expressions (same operands, same order, IEEE binary64; this
// translation unit is compiled with -fmad=false and host code never contracts).  __host__ __device__: hb_set_land folds
// on the host, hb_set_members on the device (fold_members_kernel) — one source, the same bits.
// zv = the zone's row of zpar[HB_ZP_COUNT], ev = the element's row of epar[HB_EP_COUNT]; o[j * stride] = constant j.
__host__ __device__ inline void fold_zone(const double* zv, const double* ev, int model, double* o, size_t stride) {
  const double z = ev[HB_EP_Z], relsoil = ev[HB_EP_RELSOILAREA], relupper = ev[HB_EP_RELUPPERZONEAREA],
               rellower = ev[HB_EP_RELLOWERZONEAREA], petalt = ev[HB_EP_PETALTITUDE];
  const double dz = zv[HB_ZP_ZONEZ] - z;
  const double half = zv[HB_ZP_TTINT] / 2.0;
  o[(size_t)KZ_TCORR * stride] = zv[HB_ZP_TCORR];
  o[(size_t)KZ_TCALT_DZ * stride] = zv[HB_ZP_TCALT] * dz;
  o[(size_t)KZ_TT_HI * stride] = zv[HB_ZP_TT] + half;
  o[(size_t)KZ_TT_LO * stride] = zv[HB_ZP_TT] - half;
  o[(size_t)KZ_TTINT * stride] = zv[HB_ZP_TTINT];
  o[(size_t)KZ_PCALT_F * stride] = 1.0 + zv[HB_ZP_PCALT] * dz;
  o[(size_t)KZ_PCORR * stride] = zv[HB_ZP_PCORR];
  o[(size_t)KZ_RFCF * stride] = zv[HB_ZP_RFCF];
  o[(size_t)KZ_SFCF * stride] = zv[HB_ZP_SFCF];
  o[(size_t)KZ_ICMAX * stride] = zv[HB_ZP_ICMAX];
  o[(size_t)KZ_CFMAX * stride] = zv[HB_ZP_CFMAX];
  o[(size_t)KZ_CFVAR * stride] = zv[HB_ZP_CFVAR];
  o[(size_t)KZ_TTM * stride] = zv[HB_ZP_TTM];
  o[(size_t)KZ_CFR_CFMAX * stride] = zv[HB_ZP_CFR] * zv[HB_ZP_CFMAX];
  o[(size_t)KZ_WHC * stride] = zv[HB_ZP_WHC];
  o[(size_t)KZ_FC * stride] = zv[HB_ZP_FC];
  o[(size_t)KZ_BETA * stride] = zv[HB_ZP_BETA];
  o[(size_t)KZ_CFLUX * stride] = zv[HB_ZP_CFLUX];
  o[(size_t)KZ_W_UZ * stride] = zv[HB_ZP_RELZONEAREA] / relupper;
  o[(size_t)KZ_W_CONTRI * stride] = zv[HB_ZP_RELZONEAREA] / relsoil;
  o[(size_t)KZ_THRESH * stride] = zv[HB_ZP_SOILMOISTURELIMIT] * zv[HB_ZP_MAXSOILWATER];
  o[(size_t)KZ_EXCESSRED * stride] = zv[HB_ZP_EXCESSREDUCTION];
  o[(size_t)KZ_ATF * stride] = zv[HB_ZP_AIRTEMPERATUREFACTOR];
  o[(size_t)KZ_ETF * stride] = zv[HB_ZP_EVAPOTRANSPIRATIONFACTOR];
  o[(size_t)KZ_ALT_F * stride] = 1.0 - zv[HB_ZP_ALTITUDEFACTOR] / 100.0 * (zv[HB_ZP_HRUALTITUDE] - petalt);
  o[(size_t)KZ_NEG_PF * stride] = -zv[HB_ZP_PRECIPITATIONFACTOR];
  o[(size_t)KZ_SMAX * stride] = zv[HB_ZP_SMAX];
  o[(size_t)KZ_GMELT * stride] = zv[HB_ZP_GMELT];
  o[(size_t)KZ_GVAR * stride] = zv[HB_ZP_GVAR];
  o[(size_t)KZ_W_LZ * stride] = zv[HB_ZP_RELZONEAREA] / rellower;
  o[(size_t)KZ_RELZONEAREA * stride] = zv[HB_ZP_RELZONEAREA];
  o[(size_t)KZ_TTI * stride] = zv[HB_ZP_TEMPERATURETHRESHOLDICE];
  o[(size_t)KZ_INV_TTINT * stride] = 1.0 / zv[HB_ZP_TTINT];
  o[(size_t)KZ_INV_RFCF * stride] = 1.0 / zv[HB_ZP_RFCF];
  o[(size_t)KZ_INV_SFCF * stride] = 1.0 / zv[HB_ZP_SFCF];
  o[(size_t)KZ_INV_FC * stride] = 1.0 / zv[HB_ZP_FC];
  o[(size_t)KZ_INV_THRESH * stride] = 1.0 / (zv[HB_ZP_SOILMOISTURELIMIT] * zv[HB_ZP_MAXSOILWATER]);
  if (model == HB_MODEL_HLAND_96P) {
    // folded like the reference evaluates them: (suz-sgr)*(1-w0), suz*(1-w1), w2*sg1 + ((1-w2)*k2)*gr1
    o[(size_t)KZ_X0 * stride] = zv[HB_ZP_SGR];
    o[(size_t)KZ_X1 * stride] = 1.0 - zv[HB_ZP_W0];
    o[(size_t)KZ_X2 * stride] = 1.0 - zv[HB_ZP_W1];
    o[(size_t)KZ_X3 * stride] = zv[HB_ZP_K2];
    o[(size_t)KZ_X4 * stride] = zv[HB_ZP_W2];
    o[(size_t)KZ_X5 * stride] = (1.0 - zv[HB_ZP_W2]) * zv[HB_ZP_K2];
    o[(size_t)KZ_X6 * stride] = zv[HB_ZP_SG1MAX];
  } else if (model == HB_MODEL_HLAND_96C) {
    o[(size_t)KZ_X0 * stride] = zv[HB_ZP_H1];
    o[(size_t)KZ_X1 * stride] = zv[HB_ZP_TAB1];
    o[(size_t)KZ_X2 * stride] = zv[HB_ZP_TVS1];
    o[(size_t)KZ_X3 * stride] = zv[HB_ZP_H2];
    o[(size_t)KZ_X4 * stride] = zv[HB_ZP_TAB2];
    o[(size_t)KZ_X5 * stride] = zv[HB_ZP_TVS2];
    o[(size_t)KZ_W_LAND * stride] = zv[HB_ZP_RELZONEAREA] / ev[HB_EP_RELLANDAREA];
  }
}

// beta_last = Beta of the element's last zone (ref: hland_model.py:2233-2236 `contriarea **= beta[k]` after the loop)
__host__ __device__ inline void fold_element(const double* ev, double beta_last, double* o, size_t stride) {
  const double dt = ev[HB_EP_DT], relupper = ev[HB_EP_RELUPPERZONEAREA], rellower = ev[HB_EP_RELLOWERZONEAREA];
  o[(size_t)KE_DT * stride] = dt;
  o[(size_t)KE_DT_PERCMAX * stride] = dt * ev[HB_EP_PERCMAX];
  o[(size_t)KE_DT_K * stride] = dt * ev[HB_EP_K];
  o[(size_t)KE_ALPHA1 * stride] = 1.0 + ev[HB_EP_ALPHA];
  o[(size_t)KE_K4 * stride] = ev[HB_EP_K4];
  o[(size_t)KE_GAMMA1 * stride] = 1.0 + ev[HB_EP_GAMMA];
  o[(size_t)KE_RELUPPER * stride] = relupper;
  o[(size_t)KE_RELLOWER * stride] = rellower;
  o[(size_t)KE_UP_LOW * stride] = relupper / rellower;
  o[(size_t)KE_QFACTOR * stride] = ev[HB_EP_QFACTOR];
  o[(size_t)KE_RELSOILAREA * stride] = ev[HB_EP_RELSOILAREA];
  o[(size_t)KE_RELLANDAREA * stride] = ev[HB_EP_RELLANDAREA];
  o[(size_t)KE_BETA_LAST * stride] = beta_last;
  o[(size_t)KE_K3 * stride] = ev[HB_EP_K3];
  o[(size_t)KE_W3 * stride] = ev[HB_EP_W3];
  o[(size_t)KE_W4 * stride] = ev[HB_EP_W4];
  o[(size_t)KE_FSG * stride] = ev[HB_EP_FSG];
  o[(size_t)KE_OMFSG * stride] = 1.0 - ev[HB_EP_FSG];
  o[(size_t)KE_NASH_DT * stride] = ev[HB_EP_NASH_DT];
  o[(size_t)KE_NASH_DTKSC * stride] = ev[HB_EP_NASH_DT] * ev[HB_EP_NASH_KSC];
}

template <typename T>
static int upload(hb_engine* h, DBuf<T>& buf, const T* src, size_t n) {
  HB_CUDA(h, buf.reserve(n));
  if (n) HB_CUDA(h, cudaMemcpyAsync(buf.p, src, n * sizeof(T), cudaMemcpyHostToDevice, h->stream));
  return HB_OK;
}
template <typename T>
static int upload(hb_engine* h, DBuf<T>& buf, const std::vector<T>& v) {
  int rc = upload(h, buf, v.data(), v.size());
  if (rc) return rc;
  HB_CUDA(h, cudaStreamSynchronize(h->stream));  // the vector may die right after the call
  return HB_OK;
}

extern "C" {

int hb_abi_version(void) { return HB_ABI_VERSION; }

const char* hb_last_error(const hb_engine* h) { return h ? h->error.c_str() : g_create_error.c_str(); }

int hb_create(int device, hb_engine** out) {
  if (!out) { g_create_error = "hb_create: out is NULL"; return HB_EINVAL; }
  *out = nullptr;
  int count = 0;
  cudaError_t rc = cudaGetDeviceCount(&count);
  if (rc != cudaSuccess || count == 0) {
    g_create_error = std::string("hb_create: no CUDA device available (") + cudaGetErrorString(rc) +
                     "); the B200 engine has no CPU fallback";
    return HB_ECUDA;
  }
  if (device < 0 || device >= count) {
    g_create_error = "hb_create: device index out of range";
    return HB_EINVAL;
  }
  hb_engine* h = new hb_engine();
  h->device = device;
  if ((rc = cudaSetDevice(device)) != cudaSuccess || (rc = cudaGetDeviceProperties(&h->prop, device)) != cudaSuccess ||
      (rc = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) {
    g_create_error = std::string("hb_create: ") + cudaGetErrorString(rc);
    delete h;
    return HB_ECUDA;
  }
  if (h->prop.major < 10) {
    g_create_error = std::string("hb_create: device '") + h->prop.name + "' is sm_" + std::to_string(h->prop.major) +
                     std::to_string(h->prop.minor) + "; this library holds sm_100a code only";
    cudaStreamDestroy(h->stream);
    delete h;
    return HB_ECUDA;
  }
  for (auto& e : h->ev) cudaEventCreate(&e);
  cudaFuncSetAttribute(reach_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HB_BULK_SMEM);
#define HB_FUSE_ATTR(zw) cudaFuncSetAttribute(zone_kernel<true, false, HB_MODEL_HLAND_96, zw>, \
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HB_FUSE_SMEM(zw))
  HB_FUSE_ATTR(1); HB_FUSE_ATTR(2); HB_FUSE_ATTR(3); HB_FUSE_ATTR(4); HB_FUSE_ATTR(6); HB_FUSE_ATTR(8); HB_FUSE_ATTR(12);
  HB_FUSE_ATTR(16);
#undef HB_FUSE_ATTR
  {
    const int cs = (int)HB_COUPLED_SMEM(HB_COUPLED_ZMAX);
    cudaFuncSetAttribute(coupled_kernel<288, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, cs);
    cudaFuncSetAttribute(coupled_kernel<448, HB_COUPLED_B448>, cudaFuncAttributeMaxDynamicSharedMemorySize, cs);
    cudaFuncSetAttribute(coupled_kernel<640, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, cs);
    cudaFuncSetAttribute(coupled_kernel<1024, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, cs);
  }
  {
    // routing is a chain of ≈(levels + chunks) short dependent launches: at the highest priority its blocks take every
    // slot the long-running land blocks free, so the chain is not starved while it overlaps the next window's land part
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    cudaStreamCreateWithPriority(&h->s_river, cudaStreamNonBlocking, hi);
    cudaStreamCreateWithPriority(&h->s_copy, cudaStreamNonBlocking, hi);
    cudaStreamCreateWithPriority(&h->s_comm, cudaStreamNonBlocking, hi);
    cudaStreamCreateWithPriority(&h->s_h2d, cudaStreamNonBlocking, hi);
    for (int i = 0; i < 2; ++i) {
      cudaEventCreateWithFlags(&h->ev_forcing_ready[i], cudaEventDisableTiming);
      cudaEventCreateWithFlags(&h->ev_forcing_read[i], cudaEventDisableTiming);
    }
    cudaEventCreateWithFlags(&h->ev_ext_ready[0], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_ext_ready[1], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_send_done[0], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_send_done[1], cudaEventDisableTiming);
  }
  for (int i = 0; i < 2; ++i) {
    cudaEventCreateWithFlags(&h->ev_land_done[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_river_done[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_copy_done[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_join[i], cudaEventDisableTiming);
  }
  for (auto& st : h->s_slab) cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&h->s_join, cudaStreamNonBlocking);
  {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    const int pe = hi + 1 <= lo ? hi + 1 : hi;       // just below the routing
    for (auto& row : h->s_elem)
      for (auto& st : row) cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, pe);
    for (int i = 0; i < HB_MAX_SLABS; ++i) {
      cudaEventCreateWithFlags(&h->ev_slab_zdone[i], cudaEventDisableTiming);
      for (auto& e : h->ev_slab_edone[i]) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    }
  }
  cudaEventCreateWithFlags(&h->ev_slab_fork, cudaEventDisableTiming);
  for (int i = 0; i < HB_MAX_SLABS; ++i) {
    cudaEventCreateWithFlags(&h->ev_slab_zone[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_slab_done[i], cudaEventDisableTiming);
  }
  if (!h->s_river || !h->s_copy || !h->ev_join[1] || !h->s_join || !h->ev_slab_done[HB_MAX_SLABS - 1]) {
    g_create_error = "hb_create: cannot create the pipeline streams/events";
    hb_destroy(h);
    return HB_ECUDA;
  }
  {
    MathTables t;
    make_math_tables(t);
    if (h->d_tables.reserve(1) != cudaSuccess ||
        cudaMemcpy(h->d_tables.p, &t, sizeof t, cudaMemcpyHostToDevice) != cudaSuccess) {
      g_create_error = "hb_create: cannot upload the math tables";
      cudaStreamDestroy(h->stream);
      delete h;
      return HB_ECUDA;
    }
    ChainTables ct;
    make_chain_tables(ct);
    if (h->d_chain_tables.reserve(1) != cudaSuccess ||
        cudaMemcpy(h->d_chain_tables.p, &ct, sizeof ct, cudaMemcpyHostToDevice) != cudaSuccess) {
      g_create_error = "hb_create: cannot upload the math tables";
      cudaStreamDestroy(h->stream);
      delete h;
      return HB_ECUDA;
    }
  }
  if (h->d_counters.reserve(HB_CNT_COUNT) != cudaSuccess ||
      cudaMemset(h->d_counters.p, 0, HB_CNT_COUNT * sizeof(unsigned long long)) != cudaSuccess) {
    g_create_error = "hb_create: cannot allocate the work counters";
    hb_destroy(h);
    return HB_ECUDA;
  }
  *out = h;
  return HB_OK;
}

void hb_destroy(hb_engine* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->comm && h->nccl.CommDestroy) h->nccl.CommDestroy(h->comm);
  cudaStreamSynchronize(h->stream);
  if (h->s_river) cudaStreamSynchronize(h->s_river);
  if (h->s_copy) cudaStreamSynchronize(h->s_copy);
  if (h->s_comm) cudaStreamSynchronize(h->s_comm);
  if (h->s_h2d) cudaStreamSynchronize(h->s_h2d);
  for (auto& st : h->s_slab)
    if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
  if (h->s_join) { cudaStreamSynchronize(h->s_join); cudaStreamDestroy(h->s_join); }
  if (h->ev_slab_fork) cudaEventDestroy(h->ev_slab_fork);
  for (auto& row : h->s_elem)
    for (auto& st : row)
      if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
  for (int i = 0; i < HB_MAX_SLABS; ++i) {
    if (h->ev_slab_zdone[i]) cudaEventDestroy(h->ev_slab_zdone[i]);
    for (auto& e : h->ev_slab_edone[i])
      if (e) cudaEventDestroy(e);
  }
  for (int i = 0; i < HB_MAX_SLABS; ++i) {
    if (h->ev_slab_zone[i]) cudaEventDestroy(h->ev_slab_zone[i]);
    if (h->ev_slab_done[i]) cudaEventDestroy(h->ev_slab_done[i]);
  }
  for (auto& e : h->ev)
    if (e) cudaEventDestroy(e);
  if (h->ev_stop) cudaEventDestroy(h->ev_stop);
  for (auto& e : h->land_marks.ev) cudaEventDestroy(e);
  for (auto& e : h->river_marks.ev) cudaEventDestroy(e);
  for (int i = 0; i < 2; ++i)
    for (cudaEvent_t e : {h->ev_land_done[i], h->ev_river_done[i], h->ev_copy_done[i], h->ev_join[i]})
      if (e) cudaEventDestroy(e);
  if (h->h_wave_err) cudaFreeHost(h->h_wave_err);
  if (h->s_river) cudaStreamDestroy(h->s_river);
  if (h->s_copy) cudaStreamDestroy(h->s_copy);
  if (h->s_comm) cudaStreamDestroy(h->s_comm);
  if (h->s_h2d) cudaStreamDestroy(h->s_h2d);
  for (int i = 0; i < 2; ++i) {
    if (h->ev_forcing_ready[i]) cudaEventDestroy(h->ev_forcing_ready[i]);
    if (h->ev_forcing_read[i]) cudaEventDestroy(h->ev_forcing_read[i]);
  }
  for (cudaEvent_t e : {h->ev_ext_ready[0], h->ev_ext_ready[1], h->ev_send_done[0], h->ev_send_done[1]})
    if (e) cudaEventDestroy(e);
  for (int i = 0; i < hb_engine::SIN_RING; ++i) {
    if (h->h_sin[i]) cudaFreeHost(h->h_sin[i]);
    if (h->ev_sin_copied[i]) cudaEventDestroy(h->ev_sin_copied[i]);
  }
  cudaStreamDestroy(h->stream);
  delete h;
}

int hb_device_info(hb_engine* h, int* sm_count, int* clock_khz, char* name, int name_len) {
  if (!h) return HB_EINVAL;
  if (sm_count) *sm_count = h->prop.multiProcessorCount;
  if (clock_khz) cudaDeviceGetAttribute(clock_khz, cudaDevAttrClockRate, h->device);
  if (name && name_len > 0) snprintf(name, (size_t)name_len, "%s", h->prop.name);
  return HB_OK;
}

int hb_host_alloc(hb_engine* h, int64_t bytes, void** out) {
  if (!h || !out || bytes < 0) return HB_EINVAL;
  HB_CUDA(h, cudaSetDevice(h->device));
  HB_CUDA(h, cudaHostAlloc(out, (size_t)(bytes ? bytes : 1), cudaHostAllocDefault));
  return HB_OK;
}

int hb_host_free(hb_engine* h, void* p) {
  if (!h) return HB_EINVAL;
  if (p) HB_CUDA(h, cudaFreeHost(p));
  return HB_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// land setup
// ---------------------------------------------------------------------------------------------------------------------
// ---- member modifications (hb_set_members): value' = value * mul + add for selected control parameters ---------------
struct MemberMods {
  int64_t members = 1, n_mod = 0;
  const int32_t* par = nullptr;     // [n_mod] HB_ZP_* or HB_MOD_EPAR + HB_EP_*
  const double* mul = nullptr;      // [members][n_mod]
  const double* add = nullptr;
  const uint8_t* mask = nullptr;    // [n_mod][n_elem of the base basin] or NULL: elements a modification applies to
  int64_t n_elem_base = 0;
  // modifications act in the listed order on the running value (rules of disjoint selections: one of them per element)
  double val(int id, double v, int64_t m, int64_t eb) const {
    for (int64_t j = 0; j < n_mod; ++j)
      if (par[j] == id && (!mask || mask[j * n_elem_base + eb])) v = v * mul[m * n_mod + j] + add[m * n_mod + j];
    return v;
  }
  double zval(int p, double base, int64_t m, int64_t eb) const { return n_mod ? val(p, base, m, eb) : base; }
  double eval(int p, double base, int64_t m, int64_t eb) const { return n_mod ? val(HB_MOD_EPAR + p, base, m, eb) : base; }
};

// Device twin of the host folding (fold_zone / fold_element are __host__ __device__): one thread per (member, zone slot of
// the base basin).  Raw parameters of the base basin + the member's modifications -> the constants the kernels read.
__global__ void __launch_bounds__(128) fold_members_kernel(
    const double* __restrict__ zpar, const double* __restrict__ epar, const int64_t* __restrict__ zone_ptr_b,
    const int32_t* __restrict__ model_b, int64_t n_elem_b, int64_t n_zone_b, int64_t members, int64_t n_mod,
    const int32_t* __restrict__ par, const double* __restrict__ mul, const double* __restrict__ add,
    const uint8_t* __restrict__ mask, int zmax, int64_t ep, double* __restrict__ kz, double* __restrict__ ke) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // (member, base element)
  if (i >= members * n_elem_b) return;
  const int64_t m = i / n_elem_b, eb = i - m * n_elem_b, e = i;
  const int64_t z0 = zone_ptr_b[eb];
  const int n = (int)(zone_ptr_b[eb + 1] - z0);
  const double* mm = mul + m * n_mod;
  const double* aa = add + m * n_mod;
  double ev[HB_EP_COUNT];
  for (int p = 0; p < HB_EP_COUNT; ++p) ev[p] = epar[(int64_t)p * n_elem_b + eb];
  for (int64_t j = 0; j < n_mod; ++j)   // in the listed order, on the running value (MemberMods::val)
    if (par[j] >= HB_MOD_EPAR && (!mask || mask[j * n_elem_b + eb])) {
      double& v = ev[par[j] - HB_MOD_EPAR];
      v = v * mm[j] + aa[j];
    }
  double beta_last = 0.0;
  for (int k = 0; k < n; ++k) {
    double zv[HB_ZP_COUNT];
    for (int p = 0; p < HB_ZP_COUNT; ++p) zv[p] = zpar[(int64_t)p * n_zone_b + z0 + k];
    for (int64_t j = 0; j < n_mod; ++j)
      if (par[j] < HB_MOD_EPAR && (!mask || mask[j * n_elem_b + eb])) {
        double& v = zv[par[j]];
        v = v * mm[j] + aa[j];
      }
    fold_zone(zv, ev, model_b[eb], kz + ((size_t)k * KZ_COUNT) * ep + e, (size_t)ep);
    if (k == n - 1) beta_last = zv[HB_ZP_BETA];
  }
  fold_element(ev, beta_last, ke + e, (size_t)ep);
}

// All land setup: `d` describes ONE basin; with mods.members > 1 every member is a copy of it (member-major: element
// e = m * n_elem + eb, zone cells m * n_zone + …) whose control parameters follow `mods`.
static int set_land_core(hb_engine* h, const hb_land_desc* d, const MemberMods& mods) {
  const int64_t M = mods.members;
  const int64_t neb = d->n_elem, ne = M * neb;
  const int64_t n_zone = M * d->n_zone, n_snow = M * d->n_snow, n_uh = M * d->n_uh;
  HB_REQUIRE(h, neb == 0 || (d->zone_ptr && d->snow_ptr && d->sfdist_ptr && d->uh_ptr && d->sred_ptr && d->recstep &&
                             d->eflags && d->forcing_col && d->epar && d->zpar && d->zonetype && d->zflags && d->zorder),
             HB_EINVAL, "hb_set_land: NULL array in description");
  HB_REQUIRE(h, M == 1 || d->n_sred == 0, HB_EINVAL, "hb_set_members: basins with snow redistribution are not supported");
  int zmax = 1, cmax = 1, umax = 1;
  for (int64_t e = 0; e < neb; ++e) {
    const int64_t nz = d->zone_ptr[e + 1] - d->zone_ptr[e], nc = d->sfdist_ptr[e + 1] - d->sfdist_ptr[e],
                  nu = d->uh_ptr[e + 1] - d->uh_ptr[e];
    HB_REQUIRE(h, nz >= 1 && nc >= 1 && nu >= 0, HB_EINVAL, "hb_set_land: element %lld has %lld zones, %lld snow classes",
               (long long)e, (long long)nz, (long long)nc);
    HB_REQUIRE(h, d->snow_ptr[e + 1] - d->snow_ptr[e] == nz * nc, HB_EINVAL,
               "hb_set_land: snow_ptr of element %lld does not match SClass*NmbZones", (long long)e);
    HB_REQUIRE(h, d->recstep[e] >= 0, HB_EINVAL, "hb_set_land: negative recstep");
    if (nz > zmax) zmax = (int)nz;
    if (nc > cmax) cmax = (int)nc;
    if (nu > umax) umax = (int)nu;
  }
  HB_REQUIRE(h, (neb == 0 ? 0 : d->zone_ptr[neb]) == d->n_zone, HB_EINVAL, "hb_set_land: zone_ptr does not end at n_zone");
  const int64_t ep = ((ne + 127) / 128) * 128 + (ne == 0 ? 128 : 0);
  // device states survive a new parameter set only if every element keeps its slots: same zones, snow cells and
  // unit-hydrograph ordinates per element (the layout is [zmax|cmax|umax][E_pad])
  std::vector<int64_t> zone_ptr(ne + 1), snow_ptr(ne + 1), uh_ptr(ne + 1);
  for (int64_t m = 0; m < M; ++m)
    for (int64_t e = 0; e < neb; ++e) {
      zone_ptr[m * neb + e] = m * d->n_zone + d->zone_ptr[e];
      snow_ptr[m * neb + e] = m * d->n_snow + d->snow_ptr[e];
      uh_ptr[m * neb + e] = m * d->n_uh + d->uh_ptr[e];
    }
  zone_ptr[ne] = n_zone; snow_ptr[ne] = n_snow; uh_ptr[ne] = n_uh;
  const bool same_structure = h->have_land && h->n_elem == ne && h->n_zone == n_zone && h->n_snow == n_snow &&
                              h->n_uh == n_uh && h->zone_ptr == zone_ptr && h->snow_ptr == snow_ptr && h->uh_ptr == uh_ptr;
  if (!same_structure) h->have_states = false;
  h->n_elem = ne; h->e_pad = ep; h->n_zone = n_zone; h->n_snow = n_snow; h->n_uh = n_uh;
  h->zmax = zmax; h->cmax = cmax; h->umax = umax;
  h->zone_ptr.swap(zone_ptr); h->snow_ptr.swap(snow_ptr); h->uh_ptr.swap(uh_ptr);
  h->outlet_node.assign(ne, -1);
  if (d->outlet_node)
    for (int64_t m = 0; m < M; ++m)
      for (int64_t e = 0; e < neb; ++e)   // (member m's nodes follow member m-1's: hb_set_members replicates the river)
        h->outlet_node[m * neb + e] = d->outlet_node[e] < 0 ? -1 : (int32_t)(d->outlet_node[e] + m * h->member_node_stride);

  std::vector<int32_t> nz(ep, 0), nc(ep, 1), nu(ep, 0), recstep(ep, 0), einfo(ep, 0), fcol(ep, 0), nash(ep, 0);
  std::vector<int32_t> model(ne, HB_MODEL_HLAND_96), siblings[2];
  for (int64_t e = 0; e < ne; ++e) {
    const int64_t eb = e % (neb ? neb : 1);
    if (d->model) model[e] = d->model[eb];
    HB_REQUIRE(h, model[e] >= HB_MODEL_HLAND_96 && model[e] <= HB_MODEL_HLAND_96C, HB_EINVAL,
               "hb_set_land: element %lld has unknown model %d", (long long)e, model[e]);
    if (model[e] != HB_MODEL_HLAND_96) siblings[model[e] - 1].push_back((int32_t)e);
    const int kind = d->rconc ? d->rconc[eb] : HB_RCONC_UH;
    HB_REQUIRE(h, kind == HB_RCONC_UH || kind == HB_RCONC_NASH, HB_EINVAL, "hb_set_land: unknown rconc kind %d", kind);
    if (kind == HB_RCONC_NASH) {
      HB_REQUIRE(h, d->nash_steps && d->nash_steps[eb] >= 0, HB_EINVAL, "hb_set_land: rconc_nash needs nash_steps >= 0");
      const double ksc = mods.eval(HB_EP_NASH_KSC, d->epar[(int64_t)HB_EP_NASH_KSC * neb + eb], e / (neb ? neb : 1), eb);
      // ref: rconc_model.py:183-184 — no storages or an infinite coefficient pass the inflow on
      nash[e] = (d->uh_ptr[eb + 1] == d->uh_ptr[eb] || std::isinf(ksc)) ? -1 : d->nash_steps[eb];
      // (NmbSteps == 0 would yield outflow 0; the kernel's 0 means rconc_uh, so that case is refused)
      HB_REQUIRE(h, nash[e] != 0, HB_EINVAL, "hb_set_land: rconc_nash with NmbSteps == 0");
    }
  }
  const bool host_fold = M == 1;     // (members: folded on the device, the big arrays never exist on the host)
  std::vector<int32_t> zinfo((size_t)zmax * ep, 0), zorder((size_t)zmax * ep, 0);
  std::vector<int64_t> zone0(ep, 0), snow0(ep, 0), uh0(ep, 0), sred_ptr(ep + 1, 0);
  std::vector<double> kz(host_fold ? (size_t)zmax * KZ_COUNT * ep : 0, 0.0), ke(host_fold ? (size_t)KE_COUNT * ep : 0, 0.0),
      sfdist((size_t)cmax * ep, 0.0), uh((size_t)umax * ep, 0.0);
  bool any_snowlimit = false;
  int64_t n_fast = 0, n_fast_sib[2] = {0, 0}, n_fast_classes = 0;
  std::vector<int32_t> coupled_list, spec_list;
  std::vector<char> soilwarp_ok((size_t)ep, 0);   // fast hland_96 element, every zone FIELD/FOREST with FC > 0
  h->any_nonsoil_warp = false;
  h->any_nonsoil_sib[0] = h->any_nonsoil_sib[1] = false;
  for (int64_t e = 0; e < ne; ++e) {
    const int64_t m = e / neb, eb = e - m * neb;
    const int64_t z0 = d->zone_ptr[eb];
    const int n = (int)(d->zone_ptr[eb + 1] - z0);
    nz[e] = n;
    nc[e] = (int)(d->sfdist_ptr[eb + 1] - d->sfdist_ptr[eb]);
    nu[e] = (int)(d->uh_ptr[eb + 1] - d->uh_ptr[eb]);
    recstep[e] = d->recstep[eb];
    fcol[e] = d->forcing_col[eb];
    zone0[e] = h->zone_ptr[e];
    snow0[e] = h->snow_ptr[e];
    uh0[e] = h->uh_ptr[e];
    sred_ptr[e] = d->sred_ptr[eb];
    double ev[HB_EP_COUNT];
    for (int p = 0; p < HB_EP_COUNT; ++p) ev[p] = mods.eval(p, d->epar[(int64_t)p * neb + eb], m, eb);
    int info = 0;
    bool soilonly = true, coupled = false, nonsoil = false;
    if ((d->eflags[eb] & HB_EF_RESPAREA) && ev[HB_EP_RELSOILAREA] > 0.0) info |= HB_EI_RESPAREA;
    double beta_last = 0.0;
    for (int k = 0; k < n; ++k) {
      const int64_t zz = z0 + k;
      const int zt = d->zonetype[zz];
      HB_REQUIRE(h, zt >= HB_FIELD && zt <= HB_SEALED, HB_EINVAL, "hb_set_land: zone %lld has invalid type %d",
                 (long long)zz, zt);
      HB_REQUIRE(h, d->zorder[zz] >= 0 && d->zorder[zz] < n, HB_EINVAL, "hb_set_land: zorder out of range");
      if (zt == HB_ILAKE) info |= HB_EI_LAKE;
      if (zt == HB_SEALED) info |= HB_EI_SEALED;
      const double smax = mods.zval(HB_ZP_SMAX, d->zpar[(int64_t)HB_ZP_SMAX * d->n_zone + zz], m, eb);
      const double cflux = mods.zval(HB_ZP_CFLUX, d->zpar[(int64_t)HB_ZP_CFLUX * d->n_zone + zz], m, eb);
      const double fc = mods.zval(HB_ZP_FC, d->zpar[(int64_t)HB_ZP_FC * d->n_zone + zz], m, eb);
      if (zt != HB_ILAKE && !std::isinf(smax)) info |= HB_EI_SNOWLIMIT;
      zinfo[(size_t)k * ep + e] = zt | (d->zflags[zz] << 8);
      zorder[(size_t)k * ep + e] = d->zorder[zz];
      if (host_fold) {
        // folded exactly as the reference evaluates the sub-expressions (same operands, same order, IEEE binary64;
        // this translation unit is compiled with -fmad=false and host code never contracts)
        double zv[HB_ZP_COUNT];
        for (int p = 0; p < HB_ZP_COUNT; ++p) zv[p] = mods.zval(p, d->zpar[(int64_t)p * d->n_zone + zz], m, eb);
        fold_zone(zv, ev, model[e], kz.data() + ((size_t)k * KZ_COUNT) * ep + e, (size_t)ep);
      }
      if (k == n - 1) beta_last = mods.zval(HB_ZP_BETA, d->zpar[(int64_t)HB_ZP_BETA * d->n_zone + zz], m, eb);
      if (zt != HB_FIELD && zt != HB_FOREST) soilonly = false;
      if ((zt == HB_FIELD || zt == HB_FOREST) && !(cflux == 0.0)) coupled = true;
      if (!((zt == HB_FIELD || zt == HB_FOREST) && fc > 0.0)) nonsoil = true;
    }
    if (model[e] != HB_MODEL_HLAND_96) {
      // sibling models have neither a contributing area nor a capillary flux: the time-split fast path always applies
      info &= ~HB_EI_RESPAREA;
      if (!(info & HB_EI_SNOWLIMIT) && nc[e] == 1) { info |= HB_EI_FAST; ++n_fast_sib[model[e] - 1]; }
    }
    else if (!(info & HB_EI_SNOWLIMIT) && !coupled) {
      // (several snow classes: the general zone step of the fast path loops over them; not while series are recorded)
      info |= HB_EI_FAST;
      ++n_fast;
      if (nc[e] > 1) { nonsoil = true; ++n_fast_classes; }
    }
    else if ((info & HB_EI_SNOWLIMIT) && !coupled && h->land_speculate) {
      // finite SMax: speculative fast path (general zone step, which watches the limit; not while series are recorded)
      info |= HB_EI_FAST | HB_EI_SPEC;
      ++n_fast;
      nonsoil = true;
      spec_list.push_back((int32_t)e);
    }
    else if (!(info & HB_EI_SNOWLIMIT) && nc[e] == 1 && zmax <= HB_COUPLED_ZMAX && nash[e] == 0) {
      info |= HB_EI_COUPLED;          // (the fused kernel convolves a unit hydrograph: rconc_nash elements stay general)
      coupled_list.push_back((int32_t)e);
    }
    if ((info & HB_EI_FAST) && model[e] == HB_MODEL_HLAND_96 && !nonsoil && n > 0) soilwarp_ok[e] = 1;
    if ((info & HB_EI_FAST) && nonsoil)
      (model[e] == HB_MODEL_HLAND_96 ? h->any_nonsoil_warp : h->any_nonsoil_sib[model[e] - 1]) = true;
    info |= model[e] << HB_EI_MODEL_SHIFT;
    if (soilonly) info |= HB_EI_SOILONLY;
    if (info & HB_EI_SNOWLIMIT) any_snowlimit = true;
    einfo[e] = info;
    if (host_fold) fold_element(ev, beta_last, ke.data() + e, (size_t)ep);
    for (int c = 0; c < nc[e]; ++c) sfdist[(size_t)c * ep + e] = d->sfdist[d->sfdist_ptr[eb] + c];
    for (int j = 0; j < nu[e]; ++j) uh[(size_t)j * ep + e] = d->uh[d->uh_ptr[eb] + j];
  }
  for (int64_t e = ne; e <= ep; ++e) sred_ptr[e] = neb ? d->sred_ptr[neb] : 0;
  // fused reduction (land_kernel_v2.cuh, zone_kernel<…, ZW>): groups of 32 consecutive elements whose zones all run the
  // branch-free step; the group's block has one warp per zone (at most HB_FUSE_ZMAX)
  std::vector<int32_t> gwarps((size_t)(ep / 32), 0);
  int64_t n_fused = 0;
  if (zmax <= HB_FUSE_ZMAX)
    for (int64_t g = 0; g * 32 < ne; ++g) {
      const int64_t lo = g * 32, hi = std::min<int64_t>(ne, lo + 32);
      bool ok = true;
      int w = 0;
      for (int64_t e = lo; e < hi; ++e) { ok = ok && soilwarp_ok[e]; w = std::max(w, (int)nz[e]); }
      if (!ok) continue;
      gwarps[g] = w;
      for (int64_t e = lo; e < hi; ++e) einfo[e] |= HB_EI_FUSED;
      n_fused += hi - lo;
    }
  h->n_fused = n_fused;
  { int rc_ = upload(h, h->d_gwarps, gwarps); if (rc_) return rc_; }
  h->any_snowlimit = any_snowlimit;
  h->n_fast = n_fast;
  h->n_fast_classes = n_fast_classes;
  h->n_fast_sib[0] = n_fast_sib[0]; h->n_fast_sib[1] = n_fast_sib[1];
  h->coupled_list = coupled_list;
  h->spec_list = spec_list;
  { int rc_ = upload(h, h->d_spec_list, spec_list); if (rc_) return rc_; }
  if (!spec_list.empty()) {
    HB_CUDA(h, h->d_spec_flag.reserve((size_t)ep));
    HB_CUDA(h, cudaMemsetAsync(h->d_spec_flag.p, 0, (size_t)ep * sizeof(int32_t), h->stream));
  }
  if (h->h_model != model) h->have_states = false;
  h->h_model = model;
  h->any_sibling = !siblings[0].empty() || !siblings[1].empty();
  for (int m = 0; m < 2; ++m) h->all_siblings[m] = siblings[m];
  h->any_nash = false;
  for (int64_t e = 0; e < ne; ++e) h->any_nash = h->any_nash || nash[e] > 0;
  { int rc_ = upload(h, h->d_nash_steps, nash); if (rc_) return rc_; }
  h->h_einfo.assign(einfo.begin(), einfo.begin() + ne);
  h->general_key = -1;
  {
    HB_REQUIRE(h, (int64_t)zmax * ep < ((int64_t)1 << 31), HB_EINVAL, "hb_set_land: too many zone slots for 32-bit indices");
    std::vector<int32_t> zone_slot((size_t)n_zone, -1), snow_slot((size_t)n_snow, -1);
    std::vector<int64_t> zone_cell0(ep + 1, n_zone), snow_cell0(ep + 1, n_snow);
    for (int64_t e = 0; e < ne; ++e) {
      zone_cell0[e] = h->zone_ptr[e];
      snow_cell0[e] = h->snow_ptr[e];
      // (several snow classes, finite SMax: the general kernel records their series)
      if (!(einfo[e] & HB_EI_FAST) || nc[e] > 1 || (einfo[e] & HB_EI_SPEC)) continue;
      for (int k = 0; k < nz[e]; ++k) {
        zone_slot[(size_t)h->zone_ptr[e] + k] = (int32_t)((int64_t)k * ep + e);
        snow_slot[(size_t)h->snow_ptr[e] + k] = (int32_t)((int64_t)k * ep + e);   // fast path: one snow class
      }
    }
    int rc2;
    if ((rc2 = upload(h, h->d_zone_slot, zone_slot)) || (rc2 = upload(h, h->d_snow_slot, snow_slot)) ||
        (rc2 = upload(h, h->d_zone_cell0, zone_cell0)) || (rc2 = upload(h, h->d_snow_cell0, snow_cell0)))
      return rc2;
  }
  h->h_nz = nz; h->h_nc = nc; h->h_nu = nu;
  int rc;
  if ((rc = upload(h, h->d_nz, nz)) || (rc = upload(h, h->d_nc, nc)) || (rc = upload(h, h->d_nu, nu)) ||
      (rc = upload(h, h->d_recstep, recstep)) || (rc = upload(h, h->d_einfo, einfo)) ||
      (rc = upload(h, h->d_fcol, fcol)) || (rc = upload(h, h->d_zinfo, zinfo)) || (rc = upload(h, h->d_zorder, zorder)) ||
      (rc = upload(h, h->d_zone0, zone0)) || (rc = upload(h, h->d_snow0, snow0)) || (rc = upload(h, h->d_uh0, uh0)) ||
      (rc = upload(h, h->d_sred_ptr, sred_ptr)) ||
      (rc = upload(h, h->d_sfdist, sfdist)) || (rc = upload(h, h->d_uh, uh)) ||
      (rc = upload(h, h->d_coupled_list, coupled_list)))
    return rc;
  if (host_fold) {
    if ((rc = upload(h, h->d_kz, kz)) || (rc = upload(h, h->d_ke, ke))) return rc;
  } else {
    // raw parameters of the base basin + the members' modifications -> folded constants, on the device
    DBuf<double> d_zpar, d_epar, d_mul, d_add;
    DBuf<int64_t> d_zptr;
    DBuf<int32_t> d_model, d_par;
    DBuf<uint8_t> d_mask;
    if (mods.mask && (rc = upload(h, d_mask, mods.mask, (size_t)(mods.n_mod * neb)))) return rc;
    std::vector<int32_t> model_b(neb, HB_MODEL_HLAND_96);
    if (d->model) model_b.assign(d->model, d->model + neb);
    const size_t nm = (size_t)(M * mods.n_mod);
    if ((rc = upload(h, d_zpar, d->zpar, (size_t)HB_ZP_COUNT * d->n_zone)) ||
        (rc = upload(h, d_epar, d->epar, (size_t)HB_EP_COUNT * neb)) || (rc = upload(h, d_zptr, d->zone_ptr, (size_t)neb + 1)) ||
        (rc = upload(h, d_model, model_b)) || (rc = upload(h, d_par, mods.par, (size_t)mods.n_mod)) ||
        (rc = upload(h, d_mul, mods.mul, nm)) || (rc = upload(h, d_add, mods.add, nm)))
      return rc;
    HB_CUDA(h, h->d_kz.reserve((size_t)zmax * KZ_COUNT * ep));
    HB_CUDA(h, h->d_ke.reserve((size_t)KE_COUNT * ep));
    HB_CUDA(h, cudaMemsetAsync(h->d_kz.p, 0, (size_t)zmax * KZ_COUNT * ep * sizeof(double), h->stream));
    HB_CUDA(h, cudaMemsetAsync(h->d_ke.p, 0, (size_t)KE_COUNT * ep * sizeof(double), h->stream));
    fold_members_kernel<<<(unsigned)((ne + 127) / 128), 128, 0, h->stream>>>(
        d_zpar.p, d_epar.p, d_zptr.p, d_model.p, neb, d->n_zone, M, mods.n_mod, d_par.p, d_mul.p, d_add.p,
        mods.mask ? d_mask.p : nullptr, zmax, ep,
        h->d_kz.p, h->d_ke.p);
    HB_CUDA(h, cudaGetLastError());
    HB_CUDA(h, cudaStreamSynchronize(h->stream));   // (the temporary buffers die here)
  }
  if ((rc = upload(h, h->d_sred_from, d->sred_from, (size_t)d->n_sred)) ||
      (rc = upload(h, h->d_sred_to, d->sred_to, (size_t)d->n_sred)) ||
      (rc = upload(h, h->d_sred_adjust, d->sred_adjust, (size_t)d->n_sred)))
    return rc;
  const size_t zs = (size_t)zmax * ep;
  HB_CUDA(h, h->d_ic.reserve(zs)); HB_CUDA(h, h->d_sm.reserve(zs));
  HB_CUDA(h, h->d_sp.reserve(zs * cmax)); HB_CUDA(h, h->d_wc.reserve(zs * cmax));
  HB_CUDA(h, h->d_uz.reserve(ep)); HB_CUDA(h, h->d_lz.reserve(ep)); HB_CUDA(h, h->d_quh.reserve((size_t)umax * ep));
  HB_CUDA(h, h->d_scr_a.reserve(zs)); HB_CUDA(h, h->d_scr_b.reserve(zs));
  if (h->any_sibling) {
    HB_CUDA(h, h->d_zx1.reserve(zs)); HB_CUDA(h, h->d_zx2.reserve(zs));
    HB_CUDA(h, h->d_ex1.reserve(ep)); HB_CUDA(h, h->d_ex2.reserve(ep));
  }
  HB_CUDA(h, cudaMemsetAsync(h->d_scr_a.p, 0, zs * sizeof(double), h->stream));
  HB_CUDA(h, cudaMemsetAsync(h->d_scr_b.p, 0, zs * sizeof(double), h->stream));
  if (any_snowlimit) {
    HB_CUDA(h, h->d_scr_sred.reserve(zs * 6));
    HB_CUDA(h, cudaMemsetAsync(h->d_scr_sred.p, 0, zs * 6 * sizeof(double), h->stream));
  }
  HB_CUDA(h, cudaStreamSynchronize(h->stream));
  h->have_land = true;
  h->have_results = false;
  return HB_OK;
}

int hb_set_land(hb_engine* h, const hb_land_desc* d) {
  if (!h) return HB_EINVAL;
  if (h->pending) { int rc_ = hb_wait(h); if (rc_) return rc_; }
  HB_REQUIRE(h, d && d->n_elem >= 0, HB_EINVAL, "hb_set_land: invalid description");
  HB_CUDA(h, cudaSetDevice(h->device));
  MemberMods mods;
  h->members = 1;
  h->member_node_stride = 0;
  if (h->river_members > 1) { h->have_river = false; h->river_members = 1; }   // (the replicated network belongs to the old land)
  int rc = set_land_core(h, d, mods);
  if (rc) return rc;
  h->base_land.keep(d);          // (hb_set_members builds the ensemble from this copy)
  return HB_OK;
}

int hb_set_land_states(hb_engine* h, const hb_land_states* st) {
  if (!h) return HB_EINVAL;
  if (h->pending) { int rc_ = hb_wait(h); if (rc_) return rc_; }
  HB_REQUIRE(h, h->have_land, HB_ESTATE, "hb_set_land_states: call hb_set_land first");
  HB_REQUIRE(h, st && (h->n_elem == 0 || (st->ic && st->sm && st->sp && st->wc && st->uz && st->lz)), HB_EINVAL,
             "hb_set_land_states: NULL state array");
  HB_REQUIRE(h, h->n_uh == 0 || st->quh, HB_EINVAL, "hb_set_land_states: quh is NULL");
  HB_CUDA(h, cudaSetDevice(h->device));
  const int64_t ep = h->e_pad, ne = h->n_elem;
  const size_t zs = (size_t)h->zmax * ep;
  std::vector<double> ic(zs, 0.0), sm(zs, 0.0), sp(zs * h->cmax, 0.0), wc(zs * h->cmax, 0.0), uz(ep, 0.0), lz(ep, 0.0),
      quh((size_t)h->umax * ep, 0.0);
  for (int64_t e = 0; e < ne; ++e) {
    const int64_t z0 = h->zone_ptr[e], s0 = h->snow_ptr[e], u0 = h->uh_ptr[e];
    const int nz = h->h_nz[e], nc = h->h_nc[e], nu = h->h_nu[e];
    for (int k = 0; k < nz; ++k) {
      ic[(size_t)k * ep + e] = st->ic[z0 + k];
      sm[(size_t)k * ep + e] = st->sm[z0 + k];
      for (int c = 0; c < nc; ++c) {
        sp[(size_t)c * zs + (size_t)k * ep + e] = st->sp[s0 + (int64_t)c * nz + k];
        wc[(size_t)c * zs + (size_t)k * ep + e] = st->wc[s0 + (int64_t)c * nz + k];
      }
    }
    uz[e] = st->uz[e];
    lz[e] = st->lz[e];
    for (int j = 0; j < nu; ++j) quh[(size_t)j * ep + e] = st->quh[u0 + j];
  }
  int rc;
  if ((rc = upload(h, h->d_ic, ic)) || (rc = upload(h, h->d_sm, sm)) || (rc = upload(h, h->d_sp, sp)) ||
      (rc = upload(h, h->d_wc, wc)) || (rc = upload(h, h->d_uz, uz)) || (rc = upload(h, h->d_lz, lz)) ||
      (rc = upload(h, h->d_quh, quh)))
    return rc;
  if (h->any_sibling) {
    // states of the sibling models: hland_96p suz, sg1 (zone), sg2, sg3 (element); hland_96c bw1, bw2 (zone)
    std::vector<double> zx1(zs, 0.0), zx2(zs, 0.0), ex1(ep, 0.0), ex2(ep, 0.0);
    for (int64_t e = 0; e < ne; ++e) {
      const int m = h->h_model[e];
      if (m == HB_MODEL_HLAND_96) continue;
      const int64_t z0 = h->zone_ptr[e];
      const double* a = m == HB_MODEL_HLAND_96P ? st->suz : st->bw1;
      const double* b = m == HB_MODEL_HLAND_96P ? st->sg1 : st->bw2;
      HB_REQUIRE(h, a && b && (m != HB_MODEL_HLAND_96P || (st->sg2 && st->sg3)), HB_EINVAL,
                 "hb_set_land_states: NULL state array of a sibling model (element %lld)", (long long)e);
      for (int k = 0; k < h->h_nz[e]; ++k) {
        zx1[(size_t)k * ep + e] = a[z0 + k];
        zx2[(size_t)k * ep + e] = b[z0 + k];
      }
      if (m == HB_MODEL_HLAND_96P) { ex1[e] = st->sg2[e]; ex2[e] = st->sg3[e]; }
    }
    if ((rc = upload(h, h->d_zx1, zx1)) || (rc = upload(h, h->d_zx2, zx2)) || (rc = upload(h, h->d_ex1, ex1)) ||
        (rc = upload(h, h->d_ex2, ex2)))
      return rc;
  }
  h->have_states = true;
  return HB_OK;
}

int hb_get_land_states(hb_engine* h, hb_land_states* st) {
  if (!h) return HB_EINVAL;
  if (h->pending) { int rc_ = hb_wait(h); if (rc_) return rc_; }
  HB_REQUIRE(h, h->have_land && h->have_states, HB_ESTATE, "hb_get_land_states: no states on the device");
  HB_REQUIRE(h, st, HB_EINVAL, "hb_get_land_states: NULL argument");
  HB_CUDA(h, cudaSetDevice(h->device));
  const int64_t ep = h->e_pad, ne = h->n_elem;
  const size_t zs = (size_t)h->zmax * ep;
  std::vector<double> ic(zs), sm(zs), sp(zs * h->cmax), wc(zs * h->cmax), uz(ep), lz(ep), quh((size_t)h->umax * ep);
#define HB_DL(vec, buf) HB_CUDA(h, cudaMemcpyAsync(vec.data(), buf.p, vec.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream))
  HB_DL(ic, h->d_ic); HB_DL(sm, h->d_sm); HB_DL(sp, h->d_sp); HB_DL(wc, h->d_wc); HB_DL(uz, h->d_uz); HB_DL(lz, h->d_lz);
  HB_DL(quh, h->d_quh);
#undef HB_DL
  HB_CUDA(h, cudaStreamSynchronize(h->stream));
  for (int64_t e = 0; e < ne; ++e) {
    const int64_t z0 = h->zone_ptr[e], s0 = h->snow_ptr[e], u0 = h->uh_ptr[e];
    const int nz = h->h_nz[e], nc = h->h_nc[e], nu = h->h_nu[e];
    for (int k = 0; k < nz; ++k) {
      if (st->ic) st->ic[z0 + k] = ic[(size_t)k * ep + e];
      if (st->sm) st->sm[z0 + k] = sm[(size_t)k * ep + e];
      for (int c = 0; c < nc; ++c) {
        if (st->sp) st->sp[s0 + (int64_t)c * nz + k] = sp[(size_t)c * zs + (size_t)k * ep + e];
        if (st->wc) st->wc[s0 + (int64_t)c * nz + k] = wc[(size_t)c * zs + (size_t)k * ep + e];
      }
    }
    if (st->uz) st->uz[e] = uz[e];
    if (st->lz) st->lz[e] = lz[e];
    if (st->quh)
      for (int j = 0; j < nu; ++j) st->quh[u0 + j] = quh[(size_t)j * ep + e];
  }
  // states of the sibling models (cells of elements of another model read 0)
  {
    std::vector<double> zx1, zx2, ex1, ex2;
    if (h->any_sibling) {
      zx1.resize(zs); zx2.resize(zs); ex1.resize(ep); ex2.resize(ep);
      HB_CUDA(h, cudaMemcpyAsync(zx1.data(), h->d_zx1.p, zs * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
      HB_CUDA(h, cudaMemcpyAsync(zx2.data(), h->d_zx2.p, zs * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
      HB_CUDA(h, cudaMemcpyAsync(ex1.data(), h->d_ex1.p, ep * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
      HB_CUDA(h, cudaMemcpyAsync(ex2.data(), h->d_ex2.p, ep * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
      HB_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    for (int64_t e = 0; e < ne; ++e) {
      const int m = h->h_model[e];
      const int64_t z0 = h->zone_ptr[e];
      for (int k = 0; k < h->h_nz[e]; ++k) {
        const double a = m == HB_MODEL_HLAND_96 ? 0.0 : zx1[(size_t)k * ep + e];
        const double b = m == HB_MODEL_HLAND_96 ? 0.0 : zx2[(size_t)k * ep + e];
        if (st->suz) st->suz[z0 + k] = m == HB_MODEL_HLAND_96P ? a : 0.0;
        if (st->sg1) st->sg1[z0 + k] = m == HB_MODEL_HLAND_96P ? b : 0.0;
        if (st->bw1) st->bw1[z0 + k] = m == HB_MODEL_HLAND_96C ? a : 0.0;
        if (st->bw2) st->bw2[z0 + k] = m == HB_MODEL_HLAND_96C ? b : 0.0;
      }
      if (st->sg2) st->sg2[e] = m == HB_MODEL_HLAND_96P ? ex1[e] : 0.0;
      if (st->sg3) st->sg3[e] = m == HB_MODEL_HLAND_96P ? ex2[e] : 0.0;
    }
  }
  return HB_OK;
}

int hb_set_option(hb_engine* h, int option, int64_t value) {
  if (!h) return HB_EINVAL;
  switch (option) {
    case HB_OPT_LAND_PATH:
      HB_REQUIRE(h, value == HB_LAND_AUTO || value == HB_LAND_GENERAL, HB_EINVAL, "hb_set_option: invalid land path");
      h->land_path = (int)value;
      return HB_OK;
    case HB_OPT_RIVER_WAVE_BLOCKS:
      HB_REQUIRE(h, value >= 1 && value <= 8, HB_EINVAL, "hb_set_option: wavefront blocks per SM must be 1..8");
      h->wave_blocks = (int)value;
      h->wave_blocks_set = true;
      return HB_OK;
    case HB_OPT_RIVER_STATE_SERIES:
      h->river_state_series = value != 0;
      if (!value) { h->d_dis_series[0].release(); h->d_dis_series[1].release(); }
      return HB_OK;
    case HB_OPT_RIVER_MCT_SERIES:
      HB_REQUIRE(h, value >= 0 && value < (1 << HB_MS_COUNT), HB_EINVAL, "hb_set_option: invalid musk_mct series mask");
      h->mct_series_mask = (unsigned)value;
      if (!value) { h->d_mct_series[0].release(); h->d_mct_series[1].release(); }
      return HB_OK;
    case HB_OPT_LAND_SLABS:
      HB_REQUIRE(h, value >= 1 && value <= HB_MAX_SLABS, HB_EINVAL, "hb_set_option: land slabs must be 1..%d", HB_MAX_SLABS);
      h->land_slabs = (int)value;
      return HB_OK;
    case HB_OPT_LAND_CHUNK:
      HB_REQUIRE(h, value >= -1, HB_EINVAL, "hb_set_option: invalid land chunk");
      h->land_chunk = value;
      return HB_OK;
    case HB_OPT_RIVER_PLAIN_MIN:
      HB_REQUIRE(h, value >= 0, HB_EINVAL, "hb_set_option: invalid minimum width of a plain routing level");
      h->river_plain_min = value;
      return HB_OK;
    case HB_OPT_LAND_FUSE:
      HB_REQUIRE(h, value == 0 || value == 1, HB_EINVAL, "hb_set_option: land fuse must be 0 or 1");
      h->land_fuse = (int)value;
      return HB_OK;
    case HB_OPT_LAND_SPECULATE:
      HB_REQUIRE(h, value == 0 || value == 1, HB_EINVAL, "hb_set_option: land speculate must be 0 or 1");
      h->land_speculate = (int)value;
      return HB_OK;
    case HB_OPT_RIVER_SWEEP:
      HB_REQUIRE(h, value == HB_RIVER_PERSISTENT || value == HB_RIVER_LAUNCHES || value == HB_RIVER_GROUPED, HB_EINVAL,
                 "hb_set_option: invalid routing sweep");
      h->river_sweep = (int)value;
      return HB_OK;
    default:
      return h->fail(HB_EINVAL, "hb_set_option: unknown option %d", option);
  }
}

int hb_select_series(hb_engine* h, int series, int on) {
  if (!h) return HB_EINVAL;
  HB_REQUIRE(h, series >= 0 && series < HB_S_COUNT, HB_EINVAL, "hb_select_series: unknown series %d", series);
  h->series_on[series] = on != 0;
  if (!on) h->d_series[series].release();
  return HB_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// river setup
// ---------------------------------------------------------------------------------------------------------------------
static int set_river_core(hb_engine* h, const hb_river_desc* r) {
  HB_REQUIRE(h, h->have_land, HB_ESTATE, "hb_set_river: call hb_set_land first");
  HB_REQUIRE(h, r && r->n_node >= 0 && r->n_reach >= 0, HB_EINVAL, "hb_set_river: invalid description");
  HB_CUDA(h, cudaSetDevice(h->device));
  const int64_t nn = r->n_node, nr = r->n_reach;
  HB_REQUIRE(h, (nn == 0 || (r->entry_ptr && r->node_mode && r->node_ext)) &&
                    (nr == 0 || (r->inlet_ptr && r->reach_outlet && r->seg_ptr && r->coef)),
             HB_EINVAL, "hb_set_river: NULL array in description");
  // validate + topological levels: node level = 0 without reach entries else 1 + max(level of entry reaches);
  // reach level = max(level of its inlet nodes)
  std::vector<int32_t> node_level(nn, -1), reach_level(nr, -1), node_pending(nn, 0), reach_pending(nr, 0);
  std::vector<std::vector<int32_t>> node_exits(nn);
  for (int64_t n = 0; n < nn; ++n) {
    for (int64_t i = r->entry_ptr[n]; i < r->entry_ptr[n + 1]; ++i) {
      const int32_t s = r->entry_src[i];
      if (s >= 0) HB_REQUIRE(h, s < h->n_elem, HB_EINVAL, "hb_set_river: node %lld entry refers to land element %d of %lld",
                             (long long)n, s, (long long)h->n_elem);
      else if (s <= HB_ENTRY_EXT)
        HB_REQUIRE(h, HB_ENTRY_EXT - s < r->n_ext, HB_EINVAL, "hb_set_river: node %lld entry refers to external row %d",
                   (long long)n, HB_ENTRY_EXT - s);
      else {
        HB_REQUIRE(h, -1 - s < nr, HB_EINVAL, "hb_set_river: node %lld entry refers to reach %d", (long long)n, -1 - s);
        node_pending[n]++;
      }
    }
    const int mode = r->node_mode[n];
    HB_REQUIRE(h, mode >= HB_NODE_NEWSIM && mode <= HB_NODE_OBSNEWSIM, HB_EINVAL, "hb_set_river: invalid node mode");
    if (mode != HB_NODE_NEWSIM)
      HB_REQUIRE(h, r->node_ext[n] >= 0 && r->node_ext[n] < r->n_ext, HB_EINVAL,
                 "hb_set_river: node %lld needs an external row", (long long)n);
  }
  for (int64_t q = 0; q < nr; ++q) {
    HB_REQUIRE(h, r->seg_ptr[q + 1] - r->seg_ptr[q] >= 1, HB_EINVAL, "hb_set_river: reach %lld without discharge state",
               (long long)q);
    reach_pending[q] = (int32_t)(r->inlet_ptr[q + 1] - r->inlet_ptr[q]);
    for (int64_t i = r->inlet_ptr[q]; i < r->inlet_ptr[q + 1]; ++i) {
      const int32_t n = r->inlet_node[i];
      HB_REQUIRE(h, n >= 0 && n < nn, HB_EINVAL, "hb_set_river: reach %lld inlet node out of range", (long long)q);
      node_exits[n].push_back((int32_t)q);
    }
    HB_REQUIRE(h, r->reach_outlet[q] >= -1 && r->reach_outlet[q] < nn, HB_EINVAL, "hb_set_river: outlet out of range");
  }
  std::vector<int32_t> nq, rq;  // work lists
  for (int64_t n = 0; n < nn; ++n)
    if (node_pending[n] == 0) { node_level[n] = 0; nq.push_back((int32_t)n); }
  for (int64_t q = 0; q < nr; ++q)
    if (reach_pending[q] == 0) { reach_level[q] = 0; rq.push_back((int32_t)q); }
  int64_t done = 0;
  size_t ni = 0, ri = 0;
  while (ni < nq.size() || ri < rq.size()) {
    while (ni < nq.size()) {
      const int32_t n = nq[ni++];
      ++done;
      for (int32_t q : node_exits[n]) {
        if (reach_level[q] < node_level[n]) reach_level[q] = node_level[n];
        if (--reach_pending[q] == 0) rq.push_back(q);
      }
    }
    while (ri < rq.size()) {
      const int32_t q = rq[ri++];
      ++done;
      const int32_t n = r->reach_outlet[q];
      if (n >= 0) {
        if (node_level[n] < reach_level[q] + 1) node_level[n] = reach_level[q] + 1;
        if (--node_pending[n] == 0) nq.push_back(n);
      }
    }
  }
  HB_REQUIRE(h, done == nn + nr, HB_ECYCLE, "hb_set_river: the network contains a cycle (%lld of %lld devices ordered)",
             (long long)done, (long long)(nn + nr));
  int32_t nlev = 0;
  for (auto l : node_level) if (l + 1 > nlev) nlev = l + 1;
  for (auto l : reach_level) if (l + 1 > nlev) nlev = l + 1;
  std::vector<int32_t> lnp(nlev + 1, 0), lrp(nlev + 1, 0), node_order(nn), reach_order(nr);
  for (auto l : node_level) lnp[l + 1]++;
  for (auto l : reach_level) lrp[l + 1]++;
  for (int l = 0; l < nlev; ++l) { lnp[l + 1] += lnp[l]; lrp[l + 1] += lrp[l]; }
  {
    std::vector<int32_t> np(lnp.begin(), lnp.end() - 1), rp(lrp.begin(), lrp.end() - 1);
    if (nlev == 0) { np.clear(); rp.clear(); }
    for (int64_t n = 0; n < nn; ++n) node_order[np[node_level[n]]++] = (int32_t)n;
    for (int64_t q = 0; q < nr; ++q) reach_order[rp[reach_level[q]]++] = (int32_t)q;
    // within a level: by number of segments, so that the lanes of a lane-per-reach warp run loops of similar length
    for (int l = 0; l < nlev; ++l)
      std::stable_sort(reach_order.begin() + lrp[l], reach_order.begin() + lrp[l + 1], [&](int32_t a, int32_t b) {
        return r->seg_ptr[a + 1] - r->seg_ptr[a] < r->seg_ptr[b + 1] - r->seg_ptr[b];
      });
  }
  // the reach that stores a node's row during the sweep = its first exit; nodes without exits are summed at the end
  std::vector<int32_t> node_owner(nn, -1), terminal;
  for (int64_t n = 0; n < nn; ++n) {
    if (node_exits[n].empty()) terminal.push_back((int32_t)n);
    else node_owner[n] = node_exits[n][0];
  }
  h->n_terminal = (int64_t)terminal.size();
  h->ext_feeds_reach = false;
  for (int64_t n = 0; n < nn; ++n) {
    if (node_exits[n].empty()) continue;
    if (r->node_mode[n] != HB_NODE_NEWSIM) h->ext_feeds_reach = true;
    for (int64_t i = r->entry_ptr[n]; i < r->entry_ptr[n + 1]; ++i)
      if (r->entry_src[i] <= HB_ENTRY_EXT) h->ext_feeds_reach = true;
  }
  // The wide upstream levels of a river tree are routed by two plain launches per level (reach_head_kernel: inflow of every
  // (reach, chunk); reach_bulk_kernel: lane = reach, the whole window): ≈0.3 warp instructions per reach-segment-step
  // against ≈100 per reach-step of the persistent wavefront's systolic tasks.  Level l qualifies while every level below
  // it does and at least `river_plain_min` of its reaches are simple (at most HB_BULK_SMAX segments, musk_classic, fed
  // by plain reaches only); level 0 always does.  Everything else forms the wavefront: sorted by level (and by segments
  // within a level), with the list of reaches each one waits for.
  std::vector<int32_t> head_all, head_bulk, wave_order, wave_lvl(nlev + 1, 0), dep_reach;
  std::vector<int32_t> plain_ptr_all{0}, plain_ptr_bulk{0};
  std::vector<int64_t> dep_ptr(nr + 1, 0);
  {
    std::vector<char> plain((size_t)nr, 0);
    auto simple = [&](int32_t q) {
      const int64_t nseg = r->seg_ptr[q + 1] - r->seg_ptr[q] - 1;
      if (nseg > HB_BULK_SMAX || (r->reach_model && r->reach_model[q] == HB_REACH_MCT)) return false;
      for (int64_t i = r->inlet_ptr[q]; i < r->inlet_ptr[q + 1]; ++i) {
        const int32_t n = r->inlet_node[i];
        for (int64_t k = r->entry_ptr[n]; k < r->entry_ptr[n + 1]; ++k) {
          const int32_t s_ = r->entry_src[k];
          if (s_ < 0 && s_ > HB_ENTRY_EXT && !plain[-1 - s_]) return false;
        }
      }
      return true;
    };
    int n_plain_levels = 0;
    for (int l = 0; l < nlev; ++l) {
      std::vector<int32_t> cand;
      for (int32_t i = lrp[l]; i < lrp[l + 1]; ++i)
        if (simple(reach_order[i])) cand.push_back(reach_order[i]);
      if (l > 0 && (int64_t)cand.size() < h->river_plain_min) break;
      for (int32_t q : cand) {
        plain[q] = 1;
        head_all.push_back(q);
        if (r->seg_ptr[q + 1] - r->seg_ptr[q] - 1 >= 1) head_bulk.push_back(q);
      }
      plain_ptr_all.push_back((int32_t)head_all.size());
      plain_ptr_bulk.push_back((int32_t)head_bulk.size());
      n_plain_levels = l + 1;
    }
    h->plain_ptr_all = plain_ptr_all; h->plain_ptr_bulk = plain_ptr_bulk;
    h->n_plain_levels = n_plain_levels;
    for (int32_t q : reach_order)
      if (!plain[q]) {
        wave_order.push_back(q);
        wave_lvl[reach_level[q] + 1]++;
      }
  }
  for (int l = 0; l < nlev; ++l) wave_lvl[l + 1] += wave_lvl[l];
  for (int64_t q = 0; q < nr; ++q) {
    for (int64_t i = r->inlet_ptr[q]; i < r->inlet_ptr[q + 1]; ++i) {
      const int32_t n = r->inlet_node[i];
      for (int64_t k = r->entry_ptr[n]; k < r->entry_ptr[n + 1]; ++k) {
        const int32_t s = r->entry_src[k];
        if (s < 0 && s > HB_ENTRY_EXT) dep_reach.push_back(-1 - s);
      }
    }
    dep_ptr[q + 1] = (int64_t)dep_reach.size();
  }
  HB_REQUIRE(h, r->n_seg < ((int64_t)1 << 31) && (int64_t)dep_reach.size() < ((int64_t)1 << 31) &&
                    r->n_entry < ((int64_t)1 << 31),
             HB_EINVAL, "hb_set_river: network too large for 32-bit offsets");
  std::vector<ReachMeta> meta((size_t)nr);
  for (int64_t q = 0; q < nr; ++q) {
    ReachMeta& m = meta[q];
    m.c0 = r->coef[0 * nr + q]; m.c1 = r->coef[1 * nr + q]; m.c2 = r->coef[2 * nr + q];
    m.g0 = (int32_t)r->seg_ptr[q];
    m.nseg = (int32_t)(r->seg_ptr[q + 1] - r->seg_ptr[q] - 1);
    m.level = reach_level[q];
    m.dep0 = (int32_t)dep_ptr[q];
    m.n_dep = (int32_t)(dep_ptr[q + 1] - dep_ptr[q]);
    m.simple_node = -1; m.entry0 = 0; m.n_entry = 0; m.owns_node = 0;
    m.pad = (r->reach_model && r->reach_model[q] == HB_REACH_MCT) ? 1 : 0;
    if (r->inlet_ptr[q + 1] - r->inlet_ptr[q] == 1) {
      const int32_t n = r->inlet_node[r->inlet_ptr[q]];
      const int64_t ne_ = r->entry_ptr[n + 1] - r->entry_ptr[n];
      if (r->node_mode[n] == HB_NODE_NEWSIM && ne_ <= 32) {
        m.simple_node = n;
        m.entry0 = (int32_t)r->entry_ptr[n];
        m.n_entry = (int32_t)ne_;
        m.owns_node = node_owner[n] == (int32_t)q ? 1 : 0;
      }
    }
  }
  // 128-byte wavefront records in wave order (layout: HB_WR_* in river_kernel.cuh)
  std::vector<int32_t> wave_rec(wave_order.size() * HB_WR_WORDS, 0);
  for (size_t pos = 0; pos < wave_order.size(); ++pos) {
    const int32_t q = wave_order[pos];
    const ReachMeta& m = meta[q];
    int32_t* w = wave_rec.data() + pos * HB_WR_WORDS;
    memcpy(w + HB_WR_C0, &m.c0, 8); memcpy(w + HB_WR_C1, &m.c1, 8); memcpy(w + HB_WR_C2, &m.c2, 8);
    w[HB_WR_REACH] = q; w[HB_WR_G0] = m.g0; w[HB_WR_NSEG] = m.nseg; w[HB_WR_LEVEL] = m.level;
    w[HB_WR_NDEP] = m.n_dep; w[HB_WR_NENTRY] = m.n_entry; w[HB_WR_NODE] = m.simple_node; w[HB_WR_OWNS] = m.owns_node;
    const bool inl = m.simple_node >= 0 && m.n_entry <= HB_WR_MAXLIST && m.n_dep <= HB_WR_MAXLIST && m.pad == 0;
    w[HB_WR_INLINE] = inl ? 1 : 0;
    if (inl) {
      for (int i = 0; i < m.n_entry; ++i) w[HB_WR_ENTRY + i] = r->entry_src[m.entry0 + i];
      for (int i = 0; i < m.n_dep; ++i) w[HB_WR_DEP + i] = dep_reach[m.dep0 + i];
    }
  }
  // musk_mct parameters and conditions (Courant/Reynolds numbers start at 0, the reference water depth at NaN = "no
  // start value yet", like freshly prepared reference objects; hb_set_river_states2 sets them)
  h->any_mct = false;
  for (int64_t q = 0; q < nr; ++q) {
    if (!(r->reach_model && r->reach_model[q] != HB_REACH_CLASSIC)) continue;
    HB_REQUIRE(h, r->reach_model[q] == HB_REACH_MCT, HB_EINVAL, "hb_set_river: reach %lld has unknown model %d",
               (long long)q, r->reach_model[q]);
    HB_REQUIRE(h, r->nmbruns && r->rpar && r->trap_ptr && (r->n_trap == 0 || r->tpar), HB_EINVAL,
               "hb_set_river: musk_mct reach %lld without parameters", (long long)q);
    HB_REQUIRE(h, r->nmbruns[q] >= 0 && r->trap_ptr[q + 1] >= r->trap_ptr[q] && r->trap_ptr[q + 1] <= r->n_trap, HB_EINVAL,
               "hb_set_river: invalid musk_mct description of reach %lld", (long long)q);
    h->any_mct = true;
  }
  h->n_trap = h->any_mct ? r->n_trap : 0;
  if (h->any_mct) {
    int rc_;
    if ((rc_ = upload(h, h->d_nmbruns, r->nmbruns, (size_t)nr)) || (rc_ = upload(h, h->d_rpar, r->rpar, (size_t)HB_RP_COUNT * nr)) ||
        (rc_ = upload(h, h->d_trap_ptr, r->trap_ptr, (size_t)nr + 1)) ||
        (rc_ = upload(h, h->d_tpar, r->tpar, (size_t)HB_TP_COUNT * r->n_trap)))
      return rc_;
    std::vector<double> zero((size_t)r->n_seg, 0.0), nan_((size_t)r->n_seg, std::nan(""));
    if ((rc_ = upload(h, h->d_mct_cn, zero)) || (rc_ = upload(h, h->d_mct_rn, zero)) || (rc_ = upload(h, h->d_mct_rwd, nan_)))
      return rc_;
  }
  if (const char* gw = getenv("HB_GROUP_MIN_WIDTH")) h->group_min_width = atoi(gw);   // development aid
  // units of the grouped wavefront: per level the simple reaches (one inlet node in newsim mode, at most HB_GRP_SMAX
  // segments, musk_classic) in groups of 32, every other reach on its own
  {
    std::vector<int32_t> unit_start{0}, unit_reach, unit_info, unit_lvl(nlev + 1, 0);
    for (int l = 0; l < nlev; ++l) {
      std::vector<int32_t> simple, other;
      // groups pay off where a level is wide (throughput); the narrow tail of a river tree is a chain of dependent
      // levels, where the systolic single-reach form has the shorter latency per level
      const bool wide = wave_lvl[l + 1] - wave_lvl[l] >= h->group_min_width;
      for (int i = wave_lvl[l]; i < wave_lvl[l + 1]; ++i) {
        const int32_t q = wave_order[i];
        const ReachMeta& m = meta[q];
        (wide && m.simple_node >= 0 && m.nseg <= HB_GRP_SMAX && m.n_entry <= HB_GRP_EMAX && m.n_dep <= HB_GRP_EMAX && m.pad == 0
             ? simple : other).push_back(q);
      }
      for (size_t i = 0; i < simple.size(); i += 32) {
        const size_t n = std::min<size_t>(32, simple.size() - i);
        unit_reach.insert(unit_reach.end(), simple.begin() + i, simple.begin() + i + n);
        unit_start.push_back((int32_t)unit_reach.size());
        unit_info.push_back(l | (1 << 30));
      }
      for (int32_t q : other) {
        unit_reach.push_back(q);
        unit_start.push_back((int32_t)unit_reach.size());
        unit_info.push_back(l);
      }
      unit_lvl[l + 1] = (int32_t)unit_info.size();
    }
    { std::vector<double> z(32, 0.0); int rcz = upload(h, h->d_zero_row, z); if (rcz) return rcz; }
    h->n_units = (int64_t)unit_info.size();
    h->unit_lvl_ptr = unit_lvl;
    h->diag_chunks_g = -1;
    int rc_;
    if ((rc_ = upload(h, h->d_unit_start, unit_start)) || (rc_ = upload(h, h->d_unit_reach, unit_reach)) ||
        (rc_ = upload(h, h->d_unit_info, unit_info)) || (rc_ = upload(h, h->d_unit_lvl_ptr, unit_lvl)))
      return rc_;
  }
  h->n_head_all = (int64_t)head_all.size(); h->n_head_bulk = (int64_t)head_bulk.size(); h->n_wave = (int64_t)wave_order.size();
  h->wave_lvl_ptr = wave_lvl;
  h->diag_chunks = -1;
  h->n_node = nn; h->n_reach = nr; h->n_seg = r->n_seg; h->n_ext = r->n_ext; h->n_levels = nlev;
  h->level_node_ptr = lnp; h->level_reach_ptr = lrp;
  int rc;
  if ((rc = upload(h, h->d_entry_ptr, r->entry_ptr, (size_t)nn + 1)) ||
      (rc = upload(h, h->d_entry_src, r->entry_src, (size_t)r->n_entry)) ||
      (rc = upload(h, h->d_node_mode, r->node_mode, (size_t)nn)) || (rc = upload(h, h->d_node_ext, r->node_ext, (size_t)nn)) ||
      (rc = upload(h, h->d_inlet_ptr, r->inlet_ptr, (size_t)nr + 1)) ||
      (rc = upload(h, h->d_inlet_node, r->inlet_node, (size_t)r->n_inlet)) ||
      (rc = upload(h, h->d_seg_ptr, r->seg_ptr, (size_t)nr + 1)) || (rc = upload(h, h->d_coef, r->coef, (size_t)3 * nr)) ||
      (rc = upload(h, h->d_node_order, node_order)) || (rc = upload(h, h->d_reach_order, reach_order)) ||
      (rc = upload(h, h->d_reach_level, reach_level)) || (rc = upload(h, h->d_node_owner, node_owner)) ||
      (rc = upload(h, h->d_terminal_nodes, terminal)) || (rc = upload(h, h->d_head_all, head_all)) ||
      (rc = upload(h, h->d_head_bulk, head_bulk)) || (rc = upload(h, h->d_wave_order, wave_order)) ||
      (rc = upload(h, h->d_wave_lvl_ptr, wave_lvl)) || (rc = upload(h, h->d_dep_ptr, dep_ptr)) ||
      (rc = upload(h, h->d_dep_reach, dep_reach)) || (rc = upload(h, h->d_reach_meta, meta)) ||
      (rc = upload(h, h->d_wave_rec, wave_rec)))
    return rc;
  HB_CUDA(h, h->d_progress.reserve((size_t)nr));
  HB_CUDA(h, h->d_wave_ctl.reserve(2));
  HB_CUDA(h, cudaMemset(h->d_wave_ctl.p, 0, 2 * sizeof(int)));
  if (!h->h_wave_err) HB_CUDA(h, cudaHostAlloc((void**)&h->h_wave_err, sizeof(int), cudaHostAllocDefault));
  *h->h_wave_err = 0;
  HB_CUDA(h, h->d_dis[0].reserve((size_t)r->n_seg));
  HB_CUDA(h, h->d_dis[1].reserve((size_t)r->n_seg));
  HB_CUDA(h, cudaMemsetAsync(h->d_dis[0].p, 0, (size_t)(r->n_seg ? r->n_seg : 1) * sizeof(double), h->stream));
  HB_CUDA(h, cudaMemsetAsync(h->d_dis[1].p, 0, (size_t)(r->n_seg ? r->n_seg : 1) * sizeof(double), h->stream));
  HB_CUDA(h, cudaStreamSynchronize(h->stream));
  h->dis_cur = 0;
  h->have_river = true;
  h->have_ext = false;
  return HB_OK;
}

int hb_set_river(hb_engine* h, const hb_river_desc* r) {
  if (!h) return HB_EINVAL;
  if (h->pending) { int rc_ = hb_wait(h); if (rc_) return rc_; }
  HB_REQUIRE(h, h->members == 1, HB_ESTATE, "hb_set_river: the land part holds ensemble members; describe the basin "
                                            "(hb_set_land, hb_set_river) first, then call hb_set_members");
  int rc = set_river_core(h, r);
  if (rc) return rc;
  h->base_river.keep(r);
  h->river_members = 1;
  return HB_OK;
}

// dst[row][m * n + i] = base[row][i] for m = 0 .. members-1 (row pitch `pitch` elements)
__global__ void tile_rows_kernel(double* __restrict__ dst, const double* __restrict__ base, int64_t rows, int64_t n,
                                 int64_t members, int64_t pitch) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * members * n) return;
  const int64_t row = i / (members * n), j = i - row * members * n;
  dst[row * pitch + j] = base[row * n + j % n];
}

// the first `n` cells of every row of a device state array ([rows][pitch_old]), i.e. member 0 / the base basin
static int grab_rows(hb_engine* h, std::vector<double>& out, const double* dev, int64_t rows, int64_t n, int64_t pitch_old) {
  out.assign((size_t)(rows * n), 0.0);
  if (rows * n == 0) return HB_OK;
  HB_CUDA(h, cudaMemcpy2D(out.data(), (size_t)n * sizeof(double), dev, (size_t)pitch_old * sizeof(double),
                          (size_t)n * sizeof(double), (size_t)rows, cudaMemcpyDeviceToHost));
  return HB_OK;
}
static int tile_rows(hb_engine* h, DBuf<double>& dst, const std::vector<double>& base, int64_t rows, int64_t n,
                     int64_t members, int64_t pitch) {
  if (rows * n == 0) return HB_OK;
  DBuf<double> tmp;
  int rc = upload(h, tmp, base.data(), base.size());
  if (rc) return rc;
  HB_CUDA(h, dst.reserve((size_t)(rows * pitch)));
  HB_CUDA(h, cudaMemsetAsync(dst.p, 0, (size_t)(rows * pitch) * sizeof(double), h->stream));
  const int64_t total = rows * members * n;
  tile_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, h->stream>>>(dst.p, tmp.p, rows, n, members, pitch);
  HB_CUDA(h, cudaGetLastError());
  HB_CUDA(h, cudaStreamSynchronize(h->stream));
  return HB_OK;
}

int hb_set_members(hb_engine* h, int64_t members, int64_t n_mod, const int32_t* mod_par, const double* mul,
                   const double* add, const uint8_t* mod_mask) {
  if (!h) return HB_EINVAL;
  if (h->pending) { int rc_ = hb_wait(h); if (rc_) return rc_; }
  HB_REQUIRE(h, h->have_land && h->base_land.valid, HB_ESTATE, "hb_set_members: call hb_set_land first");
  HB_REQUIRE(h, members >= 1 && n_mod >= 0 && (n_mod == 0 || (mod_par && mul && add)), HB_EINVAL,
             "hb_set_members: invalid argument");
  HB_CUDA(h, cudaSetDevice(h->device));
  const hb_land_desc* d = &h->base_land.d;
  MemberMods mods;
  mods.members = members; mods.n_mod = n_mod; mods.par = mod_par; mods.mul = mul; mods.add = add;
  mods.mask = n_mod ? mod_mask : nullptr; mods.n_elem_base = d->n_elem;
  for (int64_t j = 0; j < n_mod; ++j) {
    const int p = mod_par[j];
    const bool zone = p >= 0 && p < HB_ZP_COUNT, elem = p >= HB_MOD_EPAR && p < HB_MOD_EPAR + HB_EP_COUNT;
    HB_REQUIRE(h, zone || elem, HB_EINVAL, "hb_set_members: unknown parameter id %d", p);
    // structure and derived geometry stay what hb_set_land was given
    HB_REQUIRE(h, !(zone && (p == HB_ZP_SMAX || p == HB_ZP_RELZONEAREA || p == HB_ZP_HRUAREAFRACTION || p == HB_ZP_ZONEZ ||
                             p == HB_ZP_HRUALTITUDE)) &&
                      !(elem && (p - HB_MOD_EPAR == HB_EP_Z || p - HB_MOD_EPAR == HB_EP_RELSOILAREA ||
                                 p - HB_MOD_EPAR == HB_EP_RELLANDAREA || p - HB_MOD_EPAR == HB_EP_RELUPPERZONEAREA ||
                                 p - HB_MOD_EPAR == HB_EP_RELLOWERZONEAREA || p - HB_MOD_EPAR == HB_EP_DT ||
                                 p - HB_MOD_EPAR == HB_EP_QFACTOR || p - HB_MOD_EPAR == HB_EP_PETALTITUDE)),
               HB_EINVAL, "hb_set_members: parameter id %d is structural / derived from the geometry and cannot vary "
                          "between members", p);
  }
  // ---- the conditions of the base basin (member 0) on the device, before the arrays are rebuilt -----------------------
  const int64_t neb = d->n_elem, ep_old = h->e_pad;
  const int zmax_o = h->zmax, cmax_o = h->cmax, umax_o = h->umax;
  const bool had_states = h->have_states;
  std::vector<double> b_ic, b_sm, b_sp, b_wc, b_uz, b_lz, b_quh, b_zx1, b_zx2, b_ex1, b_ex2;
  const bool had_sib = h->any_sibling;
  if (had_states) {
    int rc;
    if ((rc = grab_rows(h, b_ic, h->d_ic.p, zmax_o, neb, ep_old)) || (rc = grab_rows(h, b_sm, h->d_sm.p, zmax_o, neb, ep_old)) ||
        (rc = grab_rows(h, b_sp, h->d_sp.p, (int64_t)zmax_o * cmax_o, neb, ep_old)) ||
        (rc = grab_rows(h, b_wc, h->d_wc.p, (int64_t)zmax_o * cmax_o, neb, ep_old)) ||
        (rc = grab_rows(h, b_uz, h->d_uz.p, 1, neb, ep_old)) || (rc = grab_rows(h, b_lz, h->d_lz.p, 1, neb, ep_old)) ||
        (rc = grab_rows(h, b_quh, h->d_quh.p, umax_o, neb, ep_old)))
      return rc;
    if (had_sib && ((rc = grab_rows(h, b_zx1, h->d_zx1.p, zmax_o, neb, ep_old)) ||
                    (rc = grab_rows(h, b_zx2, h->d_zx2.p, zmax_o, neb, ep_old)) ||
                    (rc = grab_rows(h, b_ex1, h->d_ex1.p, 1, neb, ep_old)) || (rc = grab_rows(h, b_ex2, h->d_ex2.p, 1, neb, ep_old))))
      return rc;
  }
  std::vector<double> b_dis, b_cn, b_rn, b_rwd;
  const bool had_river = h->have_river && h->base_river.valid;
  const int64_t nsb = had_river ? h->base_river.d.n_seg : 0;
  const bool had_mct = had_river && h->any_mct;
  if (had_river && nsb) {
    b_dis.resize((size_t)nsb);
    HB_CUDA(h, cudaMemcpy(b_dis.data(), h->d_dis[h->dis_cur].p, (size_t)nsb * sizeof(double), cudaMemcpyDeviceToHost));
    if (had_mct) {
      b_cn.resize((size_t)nsb); b_rn.resize((size_t)nsb); b_rwd.resize((size_t)nsb);
      HB_CUDA(h, cudaMemcpy(b_cn.data(), h->d_mct_cn.p, (size_t)nsb * sizeof(double), cudaMemcpyDeviceToHost));
      HB_CUDA(h, cudaMemcpy(b_rn.data(), h->d_mct_rn.p, (size_t)nsb * sizeof(double), cudaMemcpyDeviceToHost));
      HB_CUDA(h, cudaMemcpy(b_rwd.data(), h->d_mct_rwd.p, (size_t)nsb * sizeof(double), cudaMemcpyDeviceToHost));
    }
  }
  // ---- land: `members` copies, constants folded on the device --------------------------------------------------------
  h->members = members;
  h->member_node_stride = had_river ? h->base_river.d.n_node : 0;
  if (!had_river && d->outlet_node)
    for (int64_t e = 0; e < neb; ++e)
      if (d->outlet_node[e] + 1 > h->member_node_stride) h->member_node_stride = d->outlet_node[e] + 1;
  int rc = set_land_core(h, d, mods);
  if (rc) return rc;
  if (had_states) {
    const int64_t ep = h->e_pad;
    if ((rc = tile_rows(h, h->d_ic, b_ic, zmax_o, neb, members, ep)) || (rc = tile_rows(h, h->d_sm, b_sm, zmax_o, neb, members, ep)) ||
        (rc = tile_rows(h, h->d_sp, b_sp, (int64_t)zmax_o * cmax_o, neb, members, ep)) ||
        (rc = tile_rows(h, h->d_wc, b_wc, (int64_t)zmax_o * cmax_o, neb, members, ep)) ||
        (rc = tile_rows(h, h->d_uz, b_uz, 1, neb, members, ep)) || (rc = tile_rows(h, h->d_lz, b_lz, 1, neb, members, ep)) ||
        (rc = tile_rows(h, h->d_quh, b_quh, umax_o, neb, members, ep)))
      return rc;
    if (had_sib && ((rc = tile_rows(h, h->d_zx1, b_zx1, zmax_o, neb, members, ep)) ||
                    (rc = tile_rows(h, h->d_zx2, b_zx2, zmax_o, neb, members, ep)) ||
                    (rc = tile_rows(h, h->d_ex1, b_ex1, 1, neb, members, ep)) || (rc = tile_rows(h, h->d_ex2, b_ex2, 1, neb, members, ep))))
      return rc;
    h->have_states = true;
  }
  // ---- river: `members` copies of the network (member m's nodes, reaches and segments follow member m-1's) ------------
  if (had_river) {
    const hb_river_desc* r = &h->base_river.d;
    const int64_t nn = r->n_node, nr = r->n_reach, M = members;
    hb_river_desc t = *r;
    t.n_node = M * nn; t.n_reach = M * nr; t.n_entry = M * r->n_entry; t.n_inlet = M * r->n_inlet; t.n_seg = M * r->n_seg;
    t.n_trap = M * r->n_trap;
    std::vector<int64_t> entry_ptr(t.n_node + 1), inlet_ptr(t.n_reach + 1), seg_ptr(t.n_reach + 1), trap_ptr(t.n_reach + 1, 0);
    std::vector<int32_t> entry_src((size_t)t.n_entry), node_mode((size_t)t.n_node), node_ext((size_t)t.n_node),
        inlet_node((size_t)t.n_inlet), reach_outlet((size_t)t.n_reach), reach_model, nmbruns;
    std::vector<double> coef((size_t)3 * t.n_reach), rpar, tpar;
    HB_REQUIRE(h, t.n_node < ((int64_t)1 << 30) && t.n_reach < ((int64_t)1 << 30) && M * neb < ((int64_t)1 << 30), HB_EINVAL,
               "hb_set_members: too many members for 32-bit indices");
    if (r->reach_model) { reach_model.resize((size_t)t.n_reach); nmbruns.resize((size_t)t.n_reach); rpar.resize((size_t)HB_RP_COUNT * t.n_reach); }
    if (r->n_trap) tpar.resize((size_t)HB_TP_COUNT * t.n_trap);
    for (int64_t m = 0; m < M; ++m) {
      for (int64_t n = 0; n < nn; ++n) {
        entry_ptr[m * nn + n] = m * r->n_entry + r->entry_ptr[n];
        node_mode[m * nn + n] = r->node_mode[n];
        node_ext[m * nn + n] = r->node_ext[n];           // observation / boundary rows are shared by the members
      }
      for (int64_t i = 0; i < r->n_entry; ++i) {
        const int32_t v = r->entry_src[i];
        entry_src[m * r->n_entry + i] = v >= 0 ? (int32_t)(v + m * neb) : v <= HB_ENTRY_EXT ? v : (int32_t)(v - m * nr);
      }
      for (int64_t q = 0; q < nr; ++q) {
        inlet_ptr[m * nr + q] = m * r->n_inlet + r->inlet_ptr[q];
        seg_ptr[m * nr + q] = m * r->n_seg + r->seg_ptr[q];
        reach_outlet[m * nr + q] = r->reach_outlet[q] < 0 ? -1 : (int32_t)(r->reach_outlet[q] + m * nn);
        for (int j = 0; j < 3; ++j) coef[(size_t)j * t.n_reach + m * nr + q] = r->coef[(size_t)j * nr + q];
        if (r->reach_model) {
          reach_model[m * nr + q] = r->reach_model[q];
          nmbruns[m * nr + q] = r->nmbruns ? r->nmbruns[q] : 1;
          for (int j = 0; j < HB_RP_COUNT; ++j)
            rpar[(size_t)j * t.n_reach + m * nr + q] = r->rpar ? r->rpar[(size_t)j * nr + q] : 0.0;
        }
        if (r->trap_ptr) trap_ptr[m * nr + q] = m * r->n_trap + r->trap_ptr[q];
      }
      for (int64_t i = 0; i < r->n_inlet; ++i) inlet_node[m * r->n_inlet + i] = (int32_t)(r->inlet_node[i] + m * nn);
      for (int64_t i = 0; i < r->n_trap; ++i)
        for (int j = 0; j < HB_TP_COUNT; ++j) tpar[(size_t)j * t.n_trap + m * r->n_trap + i] = r->tpar[(size_t)j * r->n_trap + i];
    }
    entry_ptr[t.n_node] = t.n_entry; inlet_ptr[t.n_reach] = t.n_inlet; seg_ptr[t.n_reach] = t.n_seg; trap_ptr[t.n_reach] = t.n_trap;
    t.entry_ptr = entry_ptr.data(); t.entry_src = entry_src.data(); t.node_mode = node_mode.data(); t.node_ext = node_ext.data();
    t.inlet_ptr = inlet_ptr.data(); t.inlet_node = inlet_node.data(); t.reach_outlet = reach_outlet.data();
    t.seg_ptr = seg_ptr.data(); t.coef = coef.data();
    t.reach_model = r->reach_model ? reach_model.data() : nullptr; t.nmbruns = r->reach_model ? nmbruns.data() : nullptr;
    t.rpar = r->reach_model ? rpar.data() : nullptr; t.trap_ptr = r->trap_ptr ? trap_ptr.data() : nullptr;
    t.tpar = r->n_trap ? tpar.data() : nullptr;
    if ((rc = set_river_core(h, &t))) return rc;
    h->river_members = members;
    if (nsb) {
      // (discharge and the musk_mct conditions are [n_seg] arrays: the members' copies follow each other)
      for (DBuf<double>* dst : {&h->d_dis[h->dis_cur]})
        if ((rc = tile_rows(h, *dst, b_dis, 1, nsb, members, members * nsb))) return rc;
      if (had_mct && h->any_mct &&
          ((rc = tile_rows(h, h->d_mct_cn, b_cn, 1, nsb, members, members * nsb)) ||
           (rc = tile_rows(h, h->d_mct_rn, b_rn, 1, nsb, members, members * nsb)) ||
           (rc = tile_rows(h, h->d_mct_rwd, b_rwd, 1, nsb, members, members * nsb))))
        return rc;
    }
  }
  return HB_OK;
}

int hb_set_river_states(hb_engine* h, const double* discharge) {
  if (!h) return HB_EINVAL;
  if (h->pending) { int rc_ = hb_wait(h); if (rc_) return rc_; }
  HB_REQUIRE(h, h->have_river, HB_ESTATE, "hb_set_river_states: call hb_set_river first");
  HB_REQUIRE(h, discharge || h->n_seg == 0, HB_EINVAL, "hb_set_river_states: NULL argument");
  HB_CUDA(h, cudaSetDevice(h->device));
  if (h->n_seg) {
    HB_CUDA(h, cudaMemcpyAsync(h->d_dis[h->dis_cur].p, discharge, (size_t)h->n_seg * sizeof(double),
                               cudaMemcpyHostToDevice, h->stream));
    HB_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  return HB_OK;
}

int hb_set_river_states2(hb_engine* h, const hb_river_states* st) {
  if (!h) return HB_EINVAL;
  HB_REQUIRE(h, st, HB_EINVAL, "hb_set_river_states2: NULL argument");
  int rc = hb_set_river_states(h, st->discharge);
  if (rc) return rc;
  if (h->any_mct) {
    HB_REQUIRE(h, st->courantnumber && st->reynoldsnumber && st->referencewaterdepth, HB_EINVAL,
               "hb_set_river_states2: the network holds musk_mct reaches, their conditions are required");
    const size_t n = (size_t)h->n_seg;
    if ((rc = upload(h, h->d_mct_cn, st->courantnumber, n)) || (rc = upload(h, h->d_mct_rn, st->reynoldsnumber, n)) ||
        (rc = upload(h, h->d_mct_rwd, st->referencewaterdepth, n)))
      return rc;
    HB_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  return HB_OK;
}

int hb_get_river_states2(hb_engine* h, hb_river_states* st) {
  if (!h) return HB_EINVAL;
  HB_REQUIRE(h, st, HB_EINVAL, "hb_get_river_states2: NULL argument");
  int rc = hb_get_river_states(h, st->discharge);
  if (rc) return rc;
  const size_t bytes = (size_t)h->n_seg * sizeof(double);
  double* dst[3] = {st->courantnumber, st->reynoldsnumber, st->referencewaterdepth};
  double* src[3] = {h->d_mct_cn.p, h->d_mct_rn.p, h->d_mct_rwd.p};
  for (int i = 0; i < 3; ++i) {
    if (!dst[i] || !bytes) continue;
    if (h->any_mct) HB_CUDA(h, cudaMemcpy(dst[i], src[i], bytes, cudaMemcpyDeviceToHost));
    else memset(dst[i], 0, bytes);
  }
  return HB_OK;
}

int hb_get_river_states(hb_engine* h, double* discharge) {
  if (!h) return HB_EINVAL;
  if (h->pending) { int rc_ = hb_wait(h); if (rc_) return rc_; }
  HB_REQUIRE(h, h->have_river, HB_ESTATE, "hb_get_river_states: call hb_set_river first");
  HB_REQUIRE(h, discharge || h->n_seg == 0, HB_EINVAL, "hb_get_river_states: NULL argument");
  HB_CUDA(h, cudaSetDevice(h->device));
  if (h->n_seg) {
    HB_CUDA(h, cudaMemcpyAsync(discharge, h->d_dis[h->dis_cur].p, (size_t)h->n_seg * sizeof(double),
                               cudaMemcpyDeviceToHost, h->stream));
    HB_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  return HB_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// window
// ---------------------------------------------------------------------------------------------------------------------
static int stage_forcing(hb_engine* h, int64_t nt, int64_t n_col, const double* forcing, const int64_t* doy) {
  HB_REQUIRE(h, h->have_land, HB_ESTATE, "hb_set_forcing: call hb_set_land first");
  HB_REQUIRE(h, nt >= 0 && n_col >= (h->n_elem ? 1 : 0) && (nt == 0 || doy) && (nt * n_col == 0 || forcing), HB_EINVAL,
             "hb_set_forcing: invalid argument (nt=%lld, n_col=%lld)", (long long)nt, (long long)n_col);
  HB_CUDA(h, cudaSetDevice(h->device));
  const int f = h->fcur ^ 1;                      // the buffer the land run before the last one has read
  const size_t n = (size_t)HB_FORCING_COUNT * nt * n_col;
  // the copy runs on its own stream, behind the last land run that read this buffer, while the current one is busy
  HB_CUDA(h, cudaStreamWaitEvent(h->s_h2d, h->ev_forcing_read[f], 0));
  if (n > h->d_forcing[f].cap || (size_t)nt > h->d_sinfactor[f].cap) {
    HB_CUDA(h, cudaStreamSynchronize(h->stream));
    for (auto& st : h->s_slab) HB_CUDA(h, cudaStreamSynchronize(st));
  }
  HB_CUDA(h, h->d_forcing[f].reserve(n));
  HB_CUDA(h, h->d_sinfactor[f].reserve((size_t)nt));
  if (n) HB_CUDA(h, cudaMemcpyAsync(h->d_forcing[f].p, forcing, n * sizeof(double), cudaMemcpyHostToDevice, h->s_h2d));
  // 0.5*sin(2*pi*(doy+1)/366-1.39) with the HOST libm, evaluated exactly like the generated C of the reference
  // (ref: hydpy/models/hland/hland_model.py:1056-1058) — the seasonal factor is therefore bit-identical to the CPU path
  const int ss = (int)(h->sin_next++ % hb_engine::SIN_RING);
  if (!h->ev_sin_copied[ss]) HB_CUDA(h, cudaEventCreateWithFlags(&h->ev_sin_copied[ss], cudaEventDisableTiming));
  else HB_CUDA(h, cudaEventSynchronize(h->ev_sin_copied[ss]));   // the copy out of this slot, SIN_RING windows ago, is through
  if ((size_t)nt > h->h_sin_cap_ring[ss]) {
    if (h->h_sin[ss]) HB_CUDA(h, cudaFreeHost(h->h_sin[ss]));
    h->h_sin[ss] = nullptr;
    HB_CUDA(h, cudaHostAlloc((void**)&h->h_sin[ss], (size_t)nt * sizeof(double), cudaHostAllocDefault));
    h->h_sin_cap_ring[ss] = (size_t)nt;
  }
  const double pi = 3.141592653589793;
  for (int64_t t = 0; t < nt; ++t) h->h_sin[ss][t] = 0.5 * sin((((2.0 * pi) * (double)(doy[t] + 1)) / 366.0) - 1.39);
  if (nt) HB_CUDA(h, cudaMemcpyAsync(h->d_sinfactor[f].p, h->h_sin[ss], (size_t)nt * sizeof(double), cudaMemcpyHostToDevice,
                                     h->s_h2d));
  HB_CUDA(h, cudaEventRecord(h->ev_sin_copied[ss], h->s_h2d));
  HB_CUDA(h, cudaEventRecord(h->ev_forcing_ready[f], h->s_h2d));
  HB_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_forcing_ready[f], 0));   // land kernels of the new window wait for it
  h->fcur = f;
  h->nt = nt;
  h->nt_total = nt;
  h->t0 = 0;
  h->n_col = n_col;
  h->cur ^= 1;
  h->comm_calls[h->cur] = 0;
  h->have_forcing = true;
  h->have_results = false;
  h->have_routed = false;
  h->have_ext = false;
  return HB_OK;
}

int hb_set_forcing(hb_engine* h, int64_t nt, int64_t n_col, const double* forcing, const int64_t* doy) {
  if (!h) return HB_EINVAL;
  int rc = stage_forcing(h, nt, n_col, forcing, doy);
  if (rc) return rc;
  HB_CUDA(h, cudaStreamSynchronize(h->s_h2d));   // the caller's buffer is free again
  return HB_OK;
}

int hb_set_forcing_async(hb_engine* h, int64_t nt, int64_t n_col, const double* forcing, const int64_t* doy) {
  if (!h) return HB_EINVAL;
  int rc = stage_forcing(h, nt, n_col, forcing, doy);
  if (rc) return rc;
  h->pending = true;
  return HB_OK;
}

int hb_select_window(hb_engine* h, int64_t t0, int64_t nt) {
  if (!h) return HB_EINVAL;
  HB_REQUIRE(h, h->have_forcing, HB_ESTATE, "hb_select_window: call hb_set_forcing first");
  HB_REQUIRE(h, t0 >= 0 && nt >= 0 && t0 + nt <= h->nt_total, HB_EINVAL,
             "hb_select_window: [%lld, %lld) is outside the staged block of %lld steps", (long long)t0,
             (long long)(t0 + nt), (long long)h->nt_total);
  h->t0 = t0;
  h->nt = nt;
  h->cur ^= 1;
  h->comm_calls[h->cur] = 0;
  h->have_results = false;
  h->have_routed = false;
  h->have_ext = false;
  return HB_OK;
}

int hb_set_external(hb_engine* h, int64_t nt, const double* rows) {
  if (!h) return HB_EINVAL;
  HB_REQUIRE(h, h->have_river, HB_ESTATE, "hb_set_external: call hb_set_river first");
  HB_REQUIRE(h, h->have_forcing && nt == h->nt, HB_EINVAL, "hb_set_external: nt must equal the staged window length");
  HB_REQUIRE(h, rows || h->n_ext == 0 || nt == 0, HB_EINVAL, "hb_set_external: NULL argument");
  HB_CUDA(h, cudaSetDevice(h->device));
  const size_t n = (size_t)h->n_ext * nt;
  HB_CUDA(h, h->d_ext_rows[h->cur].reserve(n));
  if (n) {  // in routing-stream order: the routing of two windows ago has finished reading this buffer set
    HB_CUDA(h, cudaMemcpyAsync(h->d_ext_rows[h->cur].p, rows, n * sizeof(double), cudaMemcpyHostToDevice, h->s_river));
    HB_CUDA(h, cudaStreamSynchronize(h->s_river));
  }
  h->have_ext = true;
  return HB_OK;
}

static int run_land(hb_engine* h) {
  const int64_t nt = h->nt, ep = h->e_pad;
  HB_CUDA(h, h->d_qt.reserve((size_t)nt * ep));
  HB_CUDA(h, h->d_land_rows[h->cur].reserve((size_t)h->n_elem * nt));
  bool any_series = false;
  for (int s = 0; s < HB_S_COUNT; ++s) any_series = any_series || h->series_on[s];
  // fast path (land_kernel_v2.cuh) for the eligible elements unless series are requested or the caller forces the
  // general kernel; everything else goes through the general kernel
  // (zone/snow series of the fast path are staged through a [zmax][33] shared-memory tile: zmax <= 128)
  const bool use_fast = h->n_fast > 0 && h->land_path == HB_LAND_AUTO && !(any_series && h->zmax > 128);
  // elements with capillary flux: fused kernel (no series recording there)
  const bool use_coupled = !h->coupled_list.empty() && h->land_path == HB_LAND_AUTO && !any_series;
  // sibling models on the fast path (their zone sequences are staged like the others')
  const bool use_fast_sib = (h->n_fast_sib[0] + h->n_fast_sib[1]) > 0 && h->land_path == HB_LAND_AUTO &&
                            !(any_series && h->zmax > 128);
  // finite SMax: speculative fast path (see HB_EI_SPEC) while no series is recorded
  const bool use_spec = use_fast && !any_series && !h->spec_list.empty();
  // the general kernel takes every element the two special paths do not take in this run
  const int general_key = (use_fast ? 1 : 0) | (use_coupled ? 2 : 0) | (use_fast_sib ? 4 : 0) |
                          ((any_series && (h->n_fast_classes > 0 || !h->spec_list.empty())) ? 8 : 0);
  if (general_key != h->general_key) {
    h->general_list.clear();
    h->sibling_list[0].clear(); h->sibling_list[1].clear();
    for (int64_t e = 0; e < h->n_elem; ++e) {
      const int info = h->h_einfo[e];
      const int m = h->h_model[e];
      if (m != HB_MODEL_HLAND_96) {   // sibling models: own instances of the general kernel
        if (!(use_fast_sib && (info & HB_EI_FAST))) h->sibling_list[m - 1].push_back((int32_t)e);
        continue;
      }
      if ((use_fast && (info & HB_EI_FAST) && !(any_series && (h->h_nc[e] > 1 || (info & HB_EI_SPEC)))) ||
          (use_coupled && (info & HB_EI_COUPLED))) continue;
      h->general_list.push_back((int32_t)e);
    }
    int rc = upload(h, h->d_general_list, h->general_list);
    if (rc) return rc;
    for (int m = 0; m < 2; ++m)
      if ((rc = upload(h, h->d_sibling_list[m], h->sibling_list[m]))) return rc;
    h->general_key = general_key;
  }
  const int64_t n_general = (int64_t)h->general_list.size();
  h->timing.n_fast = (use_fast ? h->n_fast : 0) + (use_fast_sib ? h->n_fast_sib[0] + h->n_fast_sib[1] : 0);
  h->timing.n_spec = use_spec ? (int64_t)h->spec_list.size() : 0;
  h->slabs_used = 1;
  if (h->n_elem == 0 || nt == 0) {
    HB_CUDA(h, cudaEventRecord(h->ev_slab_done[0], h->stream));
    HB_CUDA(h, cudaStreamWaitEvent(h->s_join, h->ev_slab_done[0], 0));
    return HB_OK;
  }
  for (int s = 0; s < HB_S_COUNT; ++s) {
    if (!h->series_on[s]) continue;
    const int level = HB_S_LEVEL(s);
    const int64_t width = level == 0 ? h->n_zone : level == 1 ? h->n_snow : level == 2 ? h->n_elem : h->n_uh;
    HB_CUDA(h, h->d_series[s].reserve((size_t)nt * width));
    // sequences a model does not own read NaN in the cells of its elements
    const bool partial = (s >= HB_S_FIRST_ZONE2 && s < HB_S_FIRST_ZONE3) || s == HB_S_SC ||
                         (h->any_sibling && (s == HB_S_CONTRIAREA || s == HB_S_INUZ || s == HB_S_PERC || s == HB_S_Q0 ||
                                             s == HB_S_Q1 || s == HB_S_UZ || s == HB_S_LZ));
    if (partial) HB_CUDA(h, cudaMemsetAsync(h->d_series[s].p, 0xFF, (size_t)nt * width * sizeof(double), h->stream));
  }

  if (use_fast || use_fast_sib) {
    LandV2Dev v{};
    v.n_elem = h->n_elem; v.e_pad = ep; v.zmax = h->zmax;
    v.nz = h->d_nz.p; v.nu = h->d_nu.p; v.recstep = h->d_recstep.p; v.einfo = h->d_einfo.p; v.fcol = h->d_fcol.p;
    v.nc = h->d_nc.p; v.skip_classes = any_series ? 1 : 0;
    v.zinfo = h->d_zinfo.p; v.kz = h->d_kz.p; v.ke = h->d_ke.p; v.sfdist = h->d_sfdist.p; v.uh = h->d_uh.p;
    v.ic = h->d_ic.p; v.sm = h->d_sm.p; v.sp = h->d_sp.p; v.wc = h->d_wc.p; v.uz = h->d_uz.p; v.lz = h->d_lz.p;
    v.quh = h->d_quh.p; v.n_col = h->n_col; v.nt_total = h->nt_total; v.forcing = h->d_forcing[h->fcur].p;
    v.sinfactor = h->d_sinfactor[h->fcur].p; v.tables = h->d_tables.p; v.chain_tables = h->d_chain_tables.p; v.qt = h->d_qt.p;
    v.n_zone = h->n_zone; v.n_snow = h->n_snow; v.n_uh = h->n_uh; v.uh0 = h->d_uh0.p;
    v.counters = h->d_counters.p;
    v.spec_flag = h->d_spec_flag.p;
    if (use_spec) {
      // conditions of the window's start for the elements that will have to repeat it
      const size_t zs_ = (size_t)h->zmax * ep, B = sizeof(double);
      HB_CUDA(h, h->d_snap_ic.reserve(zs_)); HB_CUDA(h, h->d_snap_sm.reserve(zs_));
      HB_CUDA(h, h->d_snap_sp.reserve(zs_ * h->cmax)); HB_CUDA(h, h->d_snap_wc.reserve(zs_ * h->cmax));
      HB_CUDA(h, h->d_snap_uz.reserve((size_t)ep)); HB_CUDA(h, h->d_snap_lz.reserve((size_t)ep));
      HB_CUDA(h, h->d_snap_quh.reserve((size_t)h->umax * ep));
      HB_CUDA(h, cudaMemcpyAsync(h->d_snap_ic.p, h->d_ic.p, zs_ * B, cudaMemcpyDeviceToDevice, h->stream));
      HB_CUDA(h, cudaMemcpyAsync(h->d_snap_sm.p, h->d_sm.p, zs_ * B, cudaMemcpyDeviceToDevice, h->stream));
      HB_CUDA(h, cudaMemcpyAsync(h->d_snap_sp.p, h->d_sp.p, zs_ * h->cmax * B, cudaMemcpyDeviceToDevice, h->stream));
      HB_CUDA(h, cudaMemcpyAsync(h->d_snap_wc.p, h->d_wc.p, zs_ * h->cmax * B, cudaMemcpyDeviceToDevice, h->stream));
      HB_CUDA(h, cudaMemcpyAsync(h->d_snap_uz.p, h->d_uz.p, (size_t)ep * B, cudaMemcpyDeviceToDevice, h->stream));
      HB_CUDA(h, cudaMemcpyAsync(h->d_snap_lz.p, h->d_lz.p, (size_t)ep * B, cudaMemcpyDeviceToDevice, h->stream));
      if (h->umax > 0)
        HB_CUDA(h, cudaMemcpyAsync(h->d_snap_quh.p, h->d_quh.p, (size_t)h->umax * ep * B, cudaMemcpyDeviceToDevice, h->stream));
      // (the flags are 0 or — left by the repetition of the previous window — 2 here)
    }
    // [storage | ordinate][thread] staging of the runoff concentration: the rconc_nash cascade of a step, the rconc_uh
    // memory for the whole launch
    v.nash_stage = 1;
    // ordinates staged per element: the basin's longest unit hydrograph (at most 32) if the blocks an SM has to hold at
    // once (every chain of the slab resident: <= 12 blocks of 64 threads) still fit its shared memory, else 12
    const int slabs_est = any_series ? 1 : (h->land_slabs > 0 ? h->land_slabs : 1);   // (series recording keeps one slab)
    const int64_t eblocks_total = (ep / 64 + slabs_est - 1) / slabs_est;
    const int sms = h->prop.multiProcessorCount > 0 ? h->prop.multiProcessorCount : 148;
    const int eblk_per_sm = (int)std::min<int64_t>(12, (eblocks_total + sms - 1) / sms);
    int uh_stage = std::max(HB_NASH_STAGE, std::min(h->umax, 32));
    if ((size_t)eblk_per_sm * (sizeof(MathTables) + sizeof(ChainTables) + (size_t)uh_stage * 64 * 8 + 1024) > 227 * 1024)
      uh_stage = HB_NASH_STAGE;
    v.uh_stage = uh_stage;
    // … and the ordinates' weights behind them when there is room for both (the L1 is what the resident blocks leave of
    // the SM's 256 KB: with the launch bounds' 12 blocks in mind the driver leaves 28 KB, which the weights of a long unit
    // hydrograph miss — Lahn ensemble: 4.1 instead of 2.5 ms; an explicit carve-out per kernel is no way out, kernels
    // with different carve-outs do not share an SM)
    // (only where the element blocks stay a small tenant of the SM: a zone block of the next slab must still fit beside them)
    v.uh_weights = (size_t)eblk_per_sm * (sizeof(MathTables) + sizeof(ChainTables) + (size_t)uh_stage * 64 * 16 + 1024) <= 96 * 1024;
    const size_t esmem = (size_t)uh_stage * 64 * sizeof(double) * (v.uh_weights ? 2 : 1);
    v.nash_steps = h->d_nash_steps.p; v.zx1 = h->d_zx1.p; v.zx2 = h->d_zx2.p; v.ex1 = h->d_ex1.p; v.ex2 = h->d_ex2.p;
    int n_staged = 0;
    for (int s = 0; s < HB_S_COUNT; ++s) n_staged += (HB_S_STAGED(s) && h->series_on[s]) ? 1 : 0;
    // fused reduction: the groups whose zone terms the zone kernel adds itself (no series recording there)
    const bool fuse = use_fast && !any_series && h->land_fuse && h->n_fused > 0;
    const bool all_fused = fuse && h->n_fused == h->n_fast && !use_fast_sib;   // then no term buffers at all
    v.fused = fuse ? 1 : 0;
    h->timing.n_fused = fuse ? h->n_fused : 0;
    v.gwarps = h->d_gwarps.p;
    // time chunks bound the two term buffers (3 GB each at most)
    const int64_t per_step = all_fused ? ep : (int64_t)h->zmax * ep;   // (all fused: [chunk][E_pad] sums instead)
    // scratch per simulated step: two term buffers + one staging buffer per recorded zone/snow series.  Budget: 6 GB, or
    // up to a quarter of the free device memory (at most 32 GB) for large ensembles, whose chunks would otherwise shrink
    // to a handful of steps per launch
    double budget = 6.0e9;
    // (cudaMemGetInfo costs up to milliseconds of host time — exposed in synchronous runs — so it is only asked when the
    // window does not fit the buffers held already)
    // (all groups fused: the hand-over buffers are the two [chunk][E_pad] sum arrays, the term buffers do not exist)
    const size_t cap_a = all_fused ? h->d_red_a.cap : h->d_term_a.cap, cap_b = all_fused ? h->d_red_b.cap : h->d_term_b.cap;
    if ((size_t)nt * per_step > cap_a || (size_t)nt * per_step > cap_b || n_staged > 0) {
      size_t free_b = 0, total_b = 0;
      if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
        const double held = 8.0 * (double)(cap_a + cap_b);   // already ours, counts as available
        const double quarter = 0.25 * ((double)free_b + held);
        if (quarter > budget) budget = quarter < 32.0e9 ? quarter : 32.0e9;
      }
    }
    else budget = 8.0 * (double)(cap_a + cap_b);
    int64_t chunk_max = (int64_t)(budget / 8.0) / (per_step * (2 + n_staged));
    if (chunk_max < 8) chunk_max = 8;
    const int64_t want = h->land_chunk >= 0 ? h->land_chunk : (h->comm ? 128 : 0);
    if (want > 0 && want < chunk_max) chunk_max = want;
    const int64_t n_chunks = (nt + chunk_max - 1) / chunk_max;
    const int64_t chunk = (nt + n_chunks - 1) / n_chunks;
    if (!all_fused) {
      HB_CUDA(h, h->d_term_a.reserve((size_t)chunk * per_step));
      HB_CUDA(h, h->d_term_b.reserve((size_t)chunk * per_step));
    }
    if (fuse) {
      HB_CUDA(h, h->d_red_a.reserve((size_t)chunk * ep));
      HB_CUDA(h, h->d_red_b.reserve((size_t)chunk * ep));
    }
    v.term_a = h->d_term_a.p; v.term_b = h->d_term_b.p;
    v.red_a = h->d_red_a.p; v.red_b = h->d_red_b.p;
    for (int s = 0; s < HB_S_COUNT; ++s) {
      v.series[s] = nullptr;
      if (!h->series_on[s]) continue;
      if (HB_S_STAGED(s)) {
        HB_CUDA(h, h->d_series_stage[s].reserve((size_t)chunk * per_step));
        v.series[s] = h->d_series_stage[s].p;
      } else v.series[s] = h->d_series[s].p;
    }
    // ---- element slabs: slab s works on the elements [slab_lo[s], slab_lo[s+1]) * 128 on a stream of its own.  The zone
    // kernel of slab s starts when the zone kernel of slab s-1 (same chunk) is through, so the slabs run staggered: the
    // element kernel of one (few warps following their sequential RecStep chains, latency-bound) shares the SMs with the
    // zone kernel of the next (throughput-bound).  Series recording keeps one slab.
    const int64_t nblk_total = ep / 128;
    int n_slabs = (any_series || n_staged) ? 1 : h->land_slabs;
    if (n_slabs > nblk_total) n_slabs = (int)nblk_total;
    if (n_slabs < 1) n_slabs = 1;
    int64_t slab_lo[HB_MAX_SLABS + 1];
    for (int sl = 0; sl <= n_slabs; ++sl) slab_lo[sl] = nblk_total * sl / n_slabs;
    cudaStream_t slab_stream[HB_MAX_SLABS];
    for (int sl = 0; sl < n_slabs; ++sl) slab_stream[sl] = sl == 0 ? h->stream : h->s_slab[sl - 1];
    if (n_slabs > 1) {
      HB_CUDA(h, cudaEventRecord(h->ev_slab_fork, h->stream));   // everything the land stream has waited for so far
      for (int sl = 1; sl < n_slabs; ++sl) HB_CUDA(h, cudaStreamWaitEvent(slab_stream[sl], h->ev_slab_fork, 0));
    }
    h->slabs_used = n_slabs;
    for (int64_t c0 = 0; c0 < nt; c0 += chunk) {
      v.nt = (nt - c0 < chunk) ? (nt - c0) : chunk;
      v.t_first = h->t0 + c0;
      v.t_out = c0;
      for (int sl = 0; sl < n_slabs; ++sl) {
        cudaStream_t st = slab_stream[sl];
        v.e_lo = slab_lo[sl] * 128;
        v.slab_blocks = (int32_t)(slab_lo[sl + 1] - slab_lo[sl]);
        if (v.slab_blocks == 0) continue;
        const int64_t e_hi = std::min<int64_t>(h->n_elem, slab_lo[sl + 1] * 128);
        if (e_hi <= v.e_lo) continue;
        if (sl == 0) h->mark(HB_K_OTHER);
        else HB_CUDA(h, cudaStreamWaitEvent(st, h->ev_slab_zone[sl - 1], 0));
        const unsigned zgrid = (unsigned)((int64_t)h->zmax * v.slab_blocks), egrid = (unsigned)((e_hi - v.e_lo + 63) / 64);
        if (fuse) {
          // one block per 32-element group of the slab, one warp per zone
          const unsigned fgrid = (unsigned)v.slab_blocks * 4u;
          const int zw = h->zmax <= 1 ? 1 : h->zmax <= 2 ? 2 : h->zmax <= 3 ? 3 : h->zmax <= 4 ? 4 : h->zmax <= 6 ? 6
                         : h->zmax <= 8 ? 8 : h->zmax <= 12 ? 12 : 16;
#define HB_FUSE_LAUNCH(zw_) zone_kernel<true, false, HB_MODEL_HLAND_96, zw_><<<fgrid, zw_ * 32, HB_FUSE_SMEM(zw_), st>>>(v)
          switch (zw) {
            case 1: HB_FUSE_LAUNCH(1); break;
            case 2: HB_FUSE_LAUNCH(2); break;
            case 3: HB_FUSE_LAUNCH(3); break;
            case 4: HB_FUSE_LAUNCH(4); break;
            case 6: HB_FUSE_LAUNCH(6); break;
            case 8: HB_FUSE_LAUNCH(8); break;
            case 12: HB_FUSE_LAUNCH(12); break;
            default: HB_FUSE_LAUNCH(16); break;
          }
#undef HB_FUSE_LAUNCH
          h->timing.land_launches += 1;
        }
        if (use_fast && !all_fused) {
          if (any_series) zone_kernel<true, true, HB_MODEL_HLAND_96><<<zgrid, 128, 0, st>>>(v);
          else zone_kernel<true, false, HB_MODEL_HLAND_96><<<zgrid, 128, 0, st>>>(v);
          if (h->any_nonsoil_warp) {
            if (any_series) zone_kernel<false, true, HB_MODEL_HLAND_96><<<zgrid, 128, 0, st>>>(v);
            else zone_kernel<false, false, HB_MODEL_HLAND_96><<<zgrid, 128, 0, st>>>(v);
            h->timing.land_launches += 1;
          }
          h->timing.land_launches += 1;
        }
        // sibling models: the same grid, instantiated for their model (threads of other models retire at once).  Their own
        // zone sequences are staged like the others'; cells of another model's elements must read NaN afterwards
        if (any_series)   // (also without sibling elements: the sequences then read NaN everywhere)
          for (int s_ = HB_S_FIRST_ZONE2; s_ < HB_S_FIRST_ELEM2; ++s_)
            if (h->series_on[s_])
              HB_CUDA(h, cudaMemsetAsync(h->d_series_stage[s_].p, 0xFF, (size_t)v.nt * per_step * sizeof(double), st));
        if (use_fast_sib && h->n_fast_sib[0] > 0) {
          if (any_series) {
            zone_kernel<true, true, HB_MODEL_HLAND_96P><<<zgrid, 128, 0, st>>>(v);
            if (h->any_nonsoil_sib[0]) zone_kernel<false, true, HB_MODEL_HLAND_96P><<<zgrid, 128, 0, st>>>(v);
          } else {
            zone_kernel<true, false, HB_MODEL_HLAND_96P><<<zgrid, 128, 0, st>>>(v);
            if (h->any_nonsoil_sib[0]) zone_kernel<false, false, HB_MODEL_HLAND_96P><<<zgrid, 128, 0, st>>>(v);
          }
          h->timing.land_launches += 1 + (h->any_nonsoil_sib[0] ? 1 : 0);
        }
        if (use_fast_sib && h->n_fast_sib[1] > 0) {
          if (any_series) {
            zone_kernel<true, true, HB_MODEL_HLAND_96C><<<zgrid, 128, 0, st>>>(v);
            if (h->any_nonsoil_sib[1]) zone_kernel<false, true, HB_MODEL_HLAND_96C><<<zgrid, 128, 0, st>>>(v);
          } else {
            zone_kernel<true, false, HB_MODEL_HLAND_96C><<<zgrid, 128, 0, st>>>(v);
            if (h->any_nonsoil_sib[1]) zone_kernel<false, false, HB_MODEL_HLAND_96C><<<zgrid, 128, 0, st>>>(v);
          }
          h->timing.land_launches += 1 + (h->any_nonsoil_sib[1] ? 1 : 0);
        }
        if (n_slabs > 1) HB_CUDA(h, cudaEventRecord(h->ev_slab_zone[sl], st));
        if (n_staged) {   // (one slab)
          dim3 sg((unsigned)((h->n_elem + 31) / 32), (unsigned)(v.nt < 64 ? v.nt : 64));
          const size_t sm_bytes = (size_t)h->zmax * 33 * sizeof(double);
          for (int s = 0; s < HB_S_COUNT; ++s) {
            if (!HB_S_STAGED(s) || !h->series_on[s]) continue;
            const bool snow = HB_S_LEVEL(s) == 1;
            series_scatter_kernel<<<sg, 256, sm_bytes, st>>>(
                h->d_series_stage[s].p, h->d_series[s].p, snow ? h->d_snow_cell0.p : h->d_zone_cell0.p,
                snow ? h->d_snow_slot.p : h->d_zone_slot.p, h->n_elem, ep, h->zmax, snow ? h->n_snow : h->n_zone, (int)v.nt, c0);
            h->timing.land_launches += 1;
          }
        }
        if (sl == 0) h->mark(HB_K_ZONE);
        // element kernels: high-priority stream of the slab (see s_elem); the slab's stream continues behind them
        HB_CUDA(h, cudaEventRecord(h->ev_slab_zdone[sl], st));
        auto elem_stream = [&](int kind) -> cudaStream_t {
          cudaStream_t se = h->s_elem[sl][kind];
          cudaStreamWaitEvent(se, h->ev_slab_zdone[sl], 0);
          return se;
        };
        auto elem_done = [&](int kind) {
          cudaEventRecord(h->ev_slab_edone[sl][kind], h->s_elem[sl][kind]);
          cudaStreamWaitEvent(st, h->ev_slab_edone[sl][kind], 0);
        };
        if (fuse) {
          element_kernel<false, HB_MODEL_HLAND_96, true><<<egrid, 64, esmem, elem_stream(0)>>>(v);
          h->timing.land_launches += 1;
          elem_done(0);
        }
        if (use_fast && !all_fused) {
          cudaStream_t se = elem_stream(1);
          if (any_series) element_kernel<true, HB_MODEL_HLAND_96><<<egrid, 64, esmem, se>>>(v);
          else element_kernel<false, HB_MODEL_HLAND_96><<<egrid, 64, esmem, se>>>(v);
          h->timing.land_launches += 1;
          elem_done(1);
        }
        if (use_fast_sib && h->n_fast_sib[0] > 0) {
          if (any_series) element_kernel<true, HB_MODEL_HLAND_96P><<<egrid, 64, esmem, elem_stream(2)>>>(v);
          else element_kernel<false, HB_MODEL_HLAND_96P><<<egrid, 64, esmem, elem_stream(2)>>>(v);
          h->timing.land_launches += 1;
          elem_done(2);
        }
        if (use_fast_sib && h->n_fast_sib[1] > 0) {
          if (any_series) element_kernel<true, HB_MODEL_HLAND_96C><<<egrid, 64, esmem, elem_stream(3)>>>(v);
          else element_kernel<false, HB_MODEL_HLAND_96C><<<egrid, 64, esmem, elem_stream(3)>>>(v);
          h->timing.land_launches += 1;
          elem_done(3);
        }
        if (sl == 0) h->mark(HB_K_ELEMENT);
      }
    }
    for (int sl = 0; sl <= n_slabs; ++sl) h->slab_elem_lo[sl] = std::min<int64_t>(h->n_elem, slab_lo[sl] * 128);
    HB_CUDA(h, cudaGetLastError());
  }
  if (use_coupled) {
    LandV2Dev v{};
    v.n_elem = h->n_elem; v.e_pad = ep; v.zmax = h->zmax;
    v.nz = h->d_nz.p; v.nu = h->d_nu.p; v.recstep = h->d_recstep.p; v.einfo = h->d_einfo.p; v.fcol = h->d_fcol.p;
    v.zinfo = h->d_zinfo.p; v.kz = h->d_kz.p; v.ke = h->d_ke.p; v.sfdist = h->d_sfdist.p; v.uh = h->d_uh.p;
    v.ic = h->d_ic.p; v.sm = h->d_sm.p; v.sp = h->d_sp.p; v.wc = h->d_wc.p; v.uz = h->d_uz.p; v.lz = h->d_lz.p;
    v.quh = h->d_quh.p; v.n_col = h->n_col; v.nt_total = h->nt_total; v.forcing = h->d_forcing[h->fcur].p;
    v.sinfactor = h->d_sinfactor[h->fcur].p; v.tables = h->d_tables.p; v.chain_tables = h->d_chain_tables.p; v.qt = h->d_qt.p;
    v.nt = nt; v.t_first = h->t0; v.t_out = 0;
    v.clist = h->d_coupled_list.p; v.n_clist = (int64_t)h->coupled_list.size();
    const int zw = h->zmax;
    const unsigned threads = (unsigned)(zw + 1) * 32, grid = (unsigned)((v.n_clist + 31) / 32);
    const size_t smem = HB_COUPLED_SMEM(zw);
    h->mark(HB_K_OTHER);
    if (threads <= 288) coupled_kernel<288, 4><<<grid, threads, smem, h->stream>>>(v);
    else if (threads <= 448) coupled_kernel<448, HB_COUPLED_B448><<<grid, threads, smem, h->stream>>>(v);
    else if (threads <= 640) coupled_kernel<640, 1><<<grid, threads, smem, h->stream>>>(v);
    else coupled_kernel<1024, 1><<<grid, threads, smem, h->stream>>>(v);
    h->mark(HB_K_GENERAL);
    HB_CUDA(h, cudaGetLastError());
    h->timing.land_launches += 1;
  }
  const int64_t n_sib[2] = {(int64_t)h->sibling_list[0].size(), (int64_t)h->sibling_list[1].size()};
  LandDev d{};
  if (n_general > 0 || n_sib[0] > 0 || n_sib[1] > 0 || use_spec) {
    d.n_elem = h->n_elem; d.e_pad = ep; d.n_zone = h->n_zone; d.n_snow = h->n_snow;
    d.zmax = h->zmax; d.cmax = h->cmax; d.umax = h->umax;
    d.n_list = n_general; d.elist = n_general == h->n_elem ? nullptr : h->d_general_list.p;
    d.nz = h->d_nz.p; d.nc = h->d_nc.p; d.nu = h->d_nu.p; d.recstep = h->d_recstep.p; d.einfo = h->d_einfo.p;
    d.fcol = h->d_fcol.p; d.zinfo = h->d_zinfo.p; d.zone0 = h->d_zone0.p; d.snow0 = h->d_snow0.p; d.uh0 = h->d_uh0.p; d.n_uh = h->n_uh;
    d.kz = h->d_kz.p; d.ke = h->d_ke.p; d.sfdist = h->d_sfdist.p; d.uh = h->d_uh.p;
    d.sred_ptr = h->d_sred_ptr.p; d.sred_from = h->d_sred_from.p; d.sred_to = h->d_sred_to.p;
    d.sred_adjust = h->d_sred_adjust.p; d.zorder = h->d_zorder.p;
    d.ic = h->d_ic.p; d.sm = h->d_sm.p; d.sp = h->d_sp.p; d.wc = h->d_wc.p; d.uz = h->d_uz.p; d.lz = h->d_lz.p;
    d.quh = h->d_quh.p; d.scr_a = h->d_scr_a.p; d.scr_b = h->d_scr_b.p; d.scr_sred = h->d_scr_sred.p;
    d.nash_steps = h->d_nash_steps.p; d.zx1 = h->d_zx1.p; d.zx2 = h->d_zx2.p; d.ex1 = h->d_ex1.p; d.ex2 = h->d_ex2.p;
    d.tables = h->d_tables.p;
    d.nt = nt; d.n_col = h->n_col; d.nt_total = h->nt_total; d.t0 = h->t0; d.forcing = h->d_forcing[h->fcur].p;
    d.sinfactor = h->d_sinfactor[h->fcur].p; d.qt = h->d_qt.p;
    for (int s = 0; s < HB_S_COUNT; ++s) d.series[s] = h->series_on[s] ? h->d_series[s].p : nullptr;
  }
  if (n_general > 0 || n_sib[0] > 0 || n_sib[1] > 0) {
    const int block = 128;
    h->mark(HB_K_OTHER);
    if (n_general > 0) {
      const int grid = (int)((n_general + block - 1) / block);
      if (any_series) land_kernel<true, HB_MODEL_HLAND_96><<<grid, block, 0, h->stream>>>(d);
      else land_kernel<false, HB_MODEL_HLAND_96><<<grid, block, 0, h->stream>>>(d);
      h->timing.land_launches += 1;
    }
    // sibling models: the same kernel instantiated for hland_96p / hland_96c on their element lists
    for (int m = 0; m < 2; ++m) {
      if (n_sib[m] == 0) continue;
      d.n_list = n_sib[m]; d.elist = h->d_sibling_list[m].p;
      const int grid = (int)((n_sib[m] + block - 1) / block);
      if (m == 0) {
        if (any_series) land_kernel<true, HB_MODEL_HLAND_96P><<<grid, block, 0, h->stream>>>(d);
        else land_kernel<false, HB_MODEL_HLAND_96P><<<grid, block, 0, h->stream>>>(d);
      } else {
        if (any_series) land_kernel<true, HB_MODEL_HLAND_96C><<<grid, block, 0, h->stream>>>(d);
        else land_kernel<false, HB_MODEL_HLAND_96C><<<grid, block, 0, h->stream>>>(d);
      }
      h->timing.land_launches += 1;
    }
    h->mark(HB_K_GENERAL);
    HB_CUDA(h, cudaGetLastError());
  }
  // qt [nt][E_pad] -> land rows [n_elem][nt].  A run of fast-path slabs only: every slab transposes its own columns on its
  // own stream and the land stream goes on to the next window without waiting for the others (s_join collects them for
  // the consumers of the rows); otherwise all slabs join the land stream first, which transposes everything.
  const int ns = (use_fast || use_fast_sib) ? h->slabs_used : 1;
  const bool slabs_only = ns > 1 && !use_coupled && n_general == 0 && n_sib[0] == 0 && n_sib[1] == 0 && !use_spec;
  if (slabs_only) {
    for (int sl = 0; sl < ns; ++sl) {
      cudaStream_t st = sl == 0 ? h->stream : h->s_slab[sl - 1];
      const int64_t e_a = h->slab_elem_lo[sl], e_b = h->slab_elem_lo[sl + 1];
      if (e_b > e_a) {
        dim3 tg((unsigned)((e_b - e_a + 31) / 32), (unsigned)((nt + 31) / 32));
        transpose_qt_kernel<<<tg, 256, 0, st>>>(h->d_qt.p, h->d_land_rows[h->cur].p, nt, e_b, ep, e_a);
        h->timing.land_launches += 1;
      }
      HB_CUDA(h, cudaEventRecord(h->ev_slab_done[sl], st));
      HB_CUDA(h, cudaStreamWaitEvent(h->s_join, h->ev_slab_done[sl], 0));
    }
  } else {
    for (int sl = 1; sl < ns; ++sl) {
      HB_CUDA(h, cudaEventRecord(h->ev_slab_done[sl], h->s_slab[sl - 1]));
      HB_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_slab_done[sl], 0));
    }
    if (use_spec) {
      // every slab is through: the elements whose snow exceeded SMax somewhere in the window start again from the
      // window's conditions and run it in the general kernel, which redistributes (results and states overwritten)
      const int64_t n_spec = (int64_t)h->spec_list.size();
      const unsigned sgrid = (unsigned)((n_spec + 127) / 128);
      SpecSnap sn{h->d_ic.p, h->d_sm.p, h->d_sp.p, h->d_wc.p, h->d_uz.p, h->d_lz.p, h->d_quh.p,
                  h->d_snap_ic.p, h->d_snap_sm.p, h->d_snap_sp.p, h->d_snap_wc.p, h->d_snap_uz.p, h->d_snap_lz.p, h->d_snap_quh.p};
      h->mark(HB_K_OTHER);
      spec_restore_kernel<<<sgrid, 128, 0, h->stream>>>(h->d_spec_list.p, n_spec, h->d_spec_flag.p, ep, h->zmax, h->d_nz.p,
                                                        h->d_nc.p, h->d_nu.p, sn, h->d_counters.p);
      d.n_list = n_spec; d.elist = h->d_spec_list.p; d.only_flagged = h->d_spec_flag.p; d.flag_out = h->d_spec_flag.p;
      land_kernel<false, HB_MODEL_HLAND_96><<<sgrid, 128, 0, h->stream>>>(d);
      h->mark(HB_K_GENERAL);
      h->timing.land_launches += 2;
    }
    dim3 tg((unsigned)((h->n_elem + 31) / 32), (unsigned)((nt + 31) / 32));
    transpose_qt_kernel<<<tg, 256, 0, h->stream>>>(h->d_qt.p, h->d_land_rows[h->cur].p, nt, h->n_elem, ep, 0);
    h->timing.land_launches += 1;
    HB_CUDA(h, cudaEventRecord(h->ev_slab_done[0], h->stream));
    HB_CUDA(h, cudaStreamWaitEvent(h->s_join, h->ev_slab_done[0], 0));
  }
  HB_CUDA(h, cudaGetLastError());
  return HB_OK;
}

static int run_river(hb_engine* h) {
  const int64_t nt = h->nt;
  HB_CUDA(h, h->d_node_rows[h->cur].reserve((size_t)h->n_node * nt));
  HB_CUDA(h, h->d_reach_in[h->cur].reserve((size_t)h->n_reach * nt));
  HB_CUDA(h, h->d_reach_out[h->cur].reserve((size_t)h->n_reach * nt));
  const int n_chunks = (int)((nt + HB_RIVER_CHUNK - 1) / HB_RIVER_CHUNK);
  const int nlev = (int)h->n_levels;
  RiverDev d{};
  d.n_node = h->n_node; d.n_reach = h->n_reach; d.nt = nt;
  d.entry_ptr = h->d_entry_ptr.p; d.entry_src = h->d_entry_src.p; d.node_mode = h->d_node_mode.p;
  d.node_ext = h->d_node_ext.p; d.inlet_ptr = h->d_inlet_ptr.p; d.inlet_node = h->d_inlet_node.p;
  d.seg_ptr = h->d_seg_ptr.p; d.coef = h->d_coef.p; d.land_rows = h->d_land_rows[h->cur].p; d.ext_rows = h->d_ext_rows[h->cur].p;
  d.node_rows = h->d_node_rows[h->cur].p; d.reach_in = h->d_reach_in[h->cur].p; d.reach_out = h->d_reach_out[h->cur].p;
  d.dis[0] = h->d_dis[0].p; d.dis[1] = h->d_dis[1].p; d.par0 = h->dis_cur; d.n_chunks = n_chunks;
  d.n_seg = h->n_seg; d.dis_series = nullptr;
  if (h->river_state_series) {
    HB_CUDA(h, h->d_dis_series[h->cur].reserve((size_t)nt * h->n_seg));
    d.dis_series = h->d_dis_series[h->cur].p;
  }
  d.mct_series = nullptr; d.mct_series_mask = 0;
  h->mct_series_run[h->cur] = 0;
  if (h->mct_series_mask && h->any_mct && nt > 0) {
    // cells of musk_classic reaches and the last point of every reach stay NaN (all-ones bytes are a NaN)
    const size_t n = (size_t)__builtin_popcount(h->mct_series_mask) * nt * h->n_seg;
    HB_CUDA(h, h->d_mct_series[h->cur].reserve(n));
    HB_CUDA(h, cudaMemsetAsync(h->d_mct_series[h->cur].p, 0xFF, n * sizeof(double), h->s_river));
    d.mct_series = h->d_mct_series[h->cur].p; d.mct_series_mask = h->mct_series_mask;
    h->mct_series_run[h->cur] = h->mct_series_mask;
  }
  d.reach_level = h->d_reach_level.p; d.node_owner = h->d_node_owner.p; d.meta = h->d_reach_meta.p;
  d.wave_rec = h->d_wave_rec.p; d.wave_order = h->d_wave_order.p; d.wave_lvl_ptr = h->d_wave_lvl_ptr.p; d.dep_ptr = h->d_dep_ptr.p;
  d.dep_reach = h->d_dep_reach.p; d.n_levels = nlev; d.n_diag = nlev + n_chunks - 1;
  d.progress = h->d_progress.p; d.counter = h->d_wave_ctl.p; d.error = h->d_wave_ctl.p + 1;
  d.pscale = h->river_sweep == HB_RIVER_GROUPED ? HB_GRP_NSUB : 1;
  d.unit_start = h->d_unit_start.p; d.unit_reach = h->d_unit_reach.p; d.unit_info = h->d_unit_info.p;
  d.unit_lvl_ptr = h->d_unit_lvl_ptr.p; d.zero_row = h->d_zero_row.p;
  d.nmbruns = h->d_nmbruns.p; d.rpar = h->d_rpar.p; d.trap_ptr = h->d_trap_ptr.p; d.tpar = h->d_tpar.p; d.n_trap = h->n_trap;
  d.mct_cn = h->d_mct_cn.p; d.mct_rn = h->d_mct_rn.p; d.mct_rwd = h->d_mct_rwd.p;
  int64_t launches = 0;
  // boundary rows received over NCCL (s_comm): the whole routing waits only if some reach reads external rows;
  // otherwise just the final sums of the outlet nodes do
  HB_CUDA(h, cudaStreamWaitEvent(h->s_river, h->ev_send_done[h->cur], 0));   // rows of two windows ago are on their way
  const bool ext_pending = h->ext_async[h->cur];
  h->ext_async[h->cur] = false;
  if (ext_pending && h->ext_feeds_reach) HB_CUDA(h, cudaStreamWaitEvent(h->s_river, h->ev_ext_ready[h->cur], 0));
  if (nt > 0 && h->n_reach > 0 && h->river_sweep == HB_RIVER_LAUNCHES) {
    // one launch per anti-diagonal of the (level, chunk) grid: launch `diag` works on the reaches of level l at chunk diag - l
    for (int diag = 0; diag < nlev + n_chunks - 1; ++diag) {
      const int l_lo = diag - n_chunks + 1 > 0 ? diag - n_chunks + 1 : 0;
      const int l_hi = diag < nlev - 1 ? diag : nlev - 1;
      const int first = h->level_reach_ptr[l_lo];
      const int n = h->level_reach_ptr[l_hi + 1] - first;
      if (n <= 0) continue;
      const int block = 128, grid = (int)(((int64_t)n * 32 + block - 1) / block);
      if (h->any_mct) reach_diag_kernel<true><<<grid, block, 0, h->s_river>>>(d, h->d_reach_order.p + first, n, diag);
      else reach_diag_kernel<false><<<grid, block, 0, h->s_river>>>(d, h->d_reach_order.p + first, n, diag);
      ++launches;
    }
  } else if (nt > 0 && h->n_reach > 0) {
    // headwater reaches in two plain launches, everything else in one persistent wavefront launch
    if (h->n_wave > 0) {
      if (h->diag_chunks != n_chunks) {
        std::vector<int32_t> ds(d.n_diag + 1, 0);
        for (int dg = 0; dg < d.n_diag; ++dg) {
          const int l_lo = dg - n_chunks + 1 > 0 ? dg - n_chunks + 1 : 0;
          const int l_hi = dg < nlev - 1 ? dg : nlev - 1;
          ds[dg + 1] = ds[dg] + (h->wave_lvl_ptr[l_hi + 1] - h->wave_lvl_ptr[l_lo]);
        }
        HB_CUDA(h, cudaStreamSynchronize(h->s_river));   // a running wavefront still reads the old table
        HB_CUDA(h, h->d_diag_start.reserve(ds.size()));
        HB_CUDA(h, cudaMemcpy(h->d_diag_start.p, ds.data(), ds.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
        h->diag_chunks = n_chunks;
      }
      d.diag_start = h->d_diag_start.p;
      if (h->river_sweep == HB_RIVER_GROUPED) {
        if (h->diag_chunks_g != n_chunks) {
          // task list of the grouped wavefront, ordered by level + HB_GRP_NSUB * chunk (see reach_wave2_kernel)
          HB_REQUIRE(h, h->n_units < (1 << 24) && n_chunks < 128, HB_EINVAL,
                     "hb_run: the grouped routing sweep packs (unit, chunk) into 32 bits; use shorter windows");
          std::vector<std::vector<int32_t>> by_key((size_t)nlev + (size_t)HB_GRP_NSUB * n_chunks);
          for (int l = 0; l < nlev; ++l)
            for (int32_t u = h->unit_lvl_ptr[l]; u < h->unit_lvl_ptr[l + 1]; ++u)
              for (int c = 0; c < n_chunks; ++c) by_key[(size_t)l + (size_t)HB_GRP_NSUB * c].push_back(u | (c << 24));
          std::vector<int32_t> tasks;
          tasks.reserve((size_t)h->n_units * n_chunks);
          for (const auto& v : by_key) tasks.insert(tasks.end(), v.begin(), v.end());
          HB_CUDA(h, cudaStreamSynchronize(h->s_river));
          HB_CUDA(h, h->d_diag_start_g.reserve(tasks.size()));
          HB_CUDA(h, cudaMemcpy(h->d_diag_start_g.p, tasks.data(), tasks.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
          h->diag_chunks_g = n_chunks;
        }
        d.task_list = h->d_diag_start_g.p;
      }
      wave_init_kernel<<<(unsigned)((h->n_reach + 255) / 256), 256, 0, h->s_river>>>(d);
      ++launches;
    }
    if (h->n_head_all > 0) {
      // level by level: the outflow rows of level l - 1 are complete when the launches of level l start
      for (int l = 0; l < h->n_plain_levels; ++l) {
        const int32_t a0 = h->plain_ptr_all[l], na = h->plain_ptr_all[l + 1] - a0;
        const int32_t b0 = h->plain_ptr_bulk[l], nb = h->plain_ptr_bulk[l + 1] - b0;
        if (na > 0) {
          const int64_t warps = (int64_t)na * n_chunks;
          reach_head_kernel<<<(unsigned)((warps * 32 + 127) / 128), 128, 0, h->s_river>>>(d, h->d_head_all.p + a0, na);
          ++launches;
        }
        if (nb > 0) {
          const int per_block = 32 * HB_BULK_WARPS;
          reach_bulk_kernel<<<(unsigned)((nb + per_block - 1) / per_block), per_block, HB_BULK_SMEM, h->s_river>>>(
              d, h->d_head_bulk.p + b0, nb);
          ++launches;
        }
      }
      if (h->n_wave > 0) {
        wave_head_done_kernel<<<(unsigned)((h->n_head_all + 255) / 256), 256, 0, h->s_river>>>(d, h->d_head_all.p,
                                                                                            (int)h->n_head_all);
        ++launches;
      }
    }
    if (h->n_wave > 0 && h->river_sweep == HB_RIVER_GROUPED) {
      const int64_t n_tasks = h->n_units * n_chunks;
      HB_REQUIRE(h, n_tasks < (int64_t)1 << 31, HB_EINVAL, "hb_run: too many routing tasks for one window; use shorter windows");
      int64_t blocks = (n_tasks + HB_GRP_WARPS - 1) / HB_GRP_WARPS;
      const int64_t cap = (int64_t)h->prop.multiProcessorCount * (h->wave_blocks_set ? h->wave_blocks : 2);
      if (blocks > cap) blocks = cap;
      if (h->any_mct) reach_wave2_kernel<true><<<(unsigned)blocks, 32 * HB_GRP_WARPS, HB_GRP_SMEM, h->s_river>>>(d, (int)n_tasks);
      else reach_wave2_kernel<false><<<(unsigned)blocks, 32 * HB_GRP_WARPS, HB_GRP_SMEM, h->s_river>>>(d, (int)n_tasks);
      HB_CUDA(h, cudaMemcpyAsync(h->h_wave_err, d.error, sizeof(int), cudaMemcpyDeviceToHost, h->s_river));
#ifdef HB_GRP_DEBUG
      {
        cudaStreamSynchronize(h->s_river);
        unsigned long long c_[8];
        cudaMemcpyFromSymbol(c_, hb_grp_dbg, sizeof c_);
        fprintf(stderr, "[grp] tasks %llu; cycles per task: meta %.0f wait %.0f state+entries %.0f inflow %.0f recurrence %.0f rows+state %.0f release %.0f\n",
                c_[7], (double)c_[0] / c_[7], (double)c_[1] / c_[7], (double)c_[2] / c_[7], (double)c_[3] / c_[7],
                (double)c_[4] / c_[7], (double)c_[5] / c_[7], (double)c_[6] / c_[7]);
        unsigned long long z_[8] = {0};
        cudaMemcpyToSymbol(hb_grp_dbg, z_, sizeof z_);
      }
#endif
      ++launches;
    } else if (h->n_wave > 0) {
      const int64_t n_tasks = h->n_wave * n_chunks;
      HB_REQUIRE(h, n_tasks < (int64_t)1 << 31, HB_EINVAL, "hb_run: too many routing tasks for one window; use shorter windows");
      int64_t blocks = (n_tasks + 3) / 4;
      const int64_t cap = (int64_t)h->prop.multiProcessorCount * h->wave_blocks;
      if (blocks > cap) blocks = cap;
      const size_t tab = (size_t)(d.n_diag + 1 + nlev + 1) * sizeof(int32_t);
      const int use_smem = tab <= 40 * 1024 ? 1 : 0;
      if (h->any_mct) reach_wave_kernel<true><<<(unsigned)blocks, 128, use_smem ? tab : 0, h->s_river>>>(d, (int)n_tasks, use_smem);
      else reach_wave_kernel<false><<<(unsigned)blocks, 128, use_smem ? tab : 0, h->s_river>>>(d, (int)n_tasks, use_smem);
      HB_CUDA(h, cudaMemcpyAsync(h->h_wave_err, d.error, sizeof(int), cudaMemcpyDeviceToHost, h->s_river));
      ++launches;
    }
  }
  if (nt > 0 && h->n_reach > 0) h->dis_cur = (h->dis_cur + n_chunks) & 1;
  if (ext_pending && !h->ext_feeds_reach) HB_CUDA(h, cudaStreamWaitEvent(h->s_river, h->ev_ext_ready[h->cur], 0));
  if (nt > 0 && h->n_terminal > 0) {
    const int block = 256, grid = (int)((h->n_terminal * 32 + block - 1) / block);
    node_sum_kernel<<<grid, block, 0, h->s_river>>>(d, h->d_terminal_nodes.p, (int)h->n_terminal);
    ++launches;
  }
  HB_CUDA(h, cudaGetLastError());
  h->timing.river_launches += launches;
  return HB_OK;
}

static int enqueue_run(hb_engine* h, int flags) {
  HB_REQUIRE(h, h->have_land, HB_ESTATE, "hb_run: call hb_set_land first");
  HB_REQUIRE(h, h->have_forcing, HB_ESTATE, "hb_run: call hb_set_forcing first");
  HB_REQUIRE(h, (flags & (HB_RUN_LAND | HB_RUN_RIVER)) != 0, HB_EINVAL, "hb_run: no work requested");
  HB_CUDA(h, cudaSetDevice(h->device));
  if (flags & HB_RUN_LAND) HB_REQUIRE(h, h->have_states, HB_ESTATE, "hb_run: call hb_set_land_states first");
  if (flags & HB_RUN_RIVER) {
    HB_REQUIRE(h, h->have_river, HB_ESTATE, "hb_run: call hb_set_river first");
    HB_REQUIRE(h, h->n_ext == 0 || h->have_ext, HB_ESTATE, "hb_run: call hb_set_external for this window first");
    HB_REQUIRE(h, (flags & HB_RUN_LAND) || h->have_results, HB_ESTATE, "hb_run: routing needs the land results first");
  }
  int rc;
  const int c = h->cur;
  h->timing.n_levels = h->n_levels;
  h->pending = true;
  if (flags & HB_RUN_LAND) {
    h->have_routed = false;
    // the routing that read this set's land rows two windows ago must be through before they are overwritten
    HB_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_river_done[c], 0));
    h->land_marks.add(HB_K_BEGIN, h->stream);
    if ((rc = run_land(h))) return rc;
    h->land_marks.add(HB_K_END, h->stream);
    // (s_join has been made to wait for every slab stream of this run)
    HB_CUDA(h, cudaEventRecord(h->ev_land_done[c], h->s_join));
    HB_CUDA(h, cudaEventRecord(h->ev_forcing_read[h->fcur], h->s_join));
    h->have_results = true;
  }
  if (flags & HB_RUN_RIVER) {
    HB_CUDA(h, cudaStreamWaitEvent(h->s_river, h->ev_land_done[c], 0));
    HB_CUDA(h, cudaStreamWaitEvent(h->s_river, h->ev_copy_done[c], 0));  // rows of two windows ago are on the host
    h->river_marks.add(HB_K_BEGIN, h->s_river);
    if ((rc = run_river(h))) return rc;
    h->river_marks.add(HB_K_END, h->s_river);
    HB_CUDA(h, cudaEventRecord(h->ev_river_done[c], h->s_river));
    h->have_routed = true;
  }
  ++h->timing.n_runs;
  return HB_OK;
}

static int resolve_marks(hb_engine* h, hb_engine::Marks& m, bool land) {
  float ms = 0.f;
  size_t begin = 0;
  static const bool debug = getenv("HB_DEBUG_MARKS") != nullptr;   // development aid: every interval between two marks
  for (size_t i = 0; i < m.n; ++i) {
    const int tag = m.tag[i];
    if (debug && i > 0 && tag != HB_K_BEGIN) {
      cudaEventElapsedTime(&ms, m.ev[i - 1], m.ev[i]);
      fprintf(stderr, "[hb marks] %s interval ending at tag %d: %.3f ms\n", land ? "land" : "river", tag, ms);
    }
    static const bool timeline = getenv("HB_DEBUG_TIMELINE") != nullptr;   // development aid: when did each window's part
    if (timeline && (tag == HB_K_BEGIN || tag == HB_K_END) && h->ev[3]) {  // begin / end, relative to hb_timer_start
      float at = 0.f;
      if (cudaEventElapsedTime(&at, h->ev[3], m.ev[i]) == cudaSuccess)
        fprintf(stderr, "[hb timeline] rank %d %s %s at %.3f ms\n", h->rank, land ? "land" : "river",
                tag == HB_K_BEGIN ? "begin" : "end", at);
      else cudaGetLastError();
    }
    if (tag == HB_K_BEGIN) { begin = i; continue; }
    if (tag == HB_K_END) {
      HB_CUDA(h, cudaEventElapsedTime(&ms, m.ev[begin], m.ev[i]));
      (land ? h->timing.land_ms : h->timing.river_ms) += ms;
      continue;
    }
    if (tag == HB_K_OTHER || i == 0) continue;
    HB_CUDA(h, cudaEventElapsedTime(&ms, m.ev[i - 1], m.ev[i]));
    if (tag == HB_K_ZONE) h->timing.zone_ms += ms;
    else if (tag == HB_K_ELEMENT) h->timing.element_ms += ms;
    else if (tag == HB_K_GENERAL) h->timing.general_ms += ms;
  }
  m.n = 0;
  return HB_OK;
}

int hb_wait(hb_engine* h) {
  if (!h) return HB_EINVAL;
  HB_CUDA(h, cudaSetDevice(h->device));
  HB_CUDA(h, cudaStreamSynchronize(h->stream));
  for (auto& st : h->s_slab) HB_CUDA(h, cudaStreamSynchronize(st));
  HB_CUDA(h, cudaStreamSynchronize(h->s_join));
  HB_CUDA(h, cudaStreamSynchronize(h->s_river));
  HB_CUDA(h, cudaStreamSynchronize(h->s_copy));
  HB_CUDA(h, cudaStreamSynchronize(h->s_comm));
  HB_CUDA(h, cudaStreamSynchronize(h->s_h2d));
  int rc;
  if ((rc = resolve_marks(h, h->land_marks, true)) || (rc = resolve_marks(h, h->river_marks, false))) return rc;
  for (size_t i = 0; i < h->comm_marks.n; ++i) {   // (HB_DEBUG_TIMELINE)
    float at = 0.f;
    static const char* what[] = {"send: routing done", "send: rows packed", "send: done", "recv: done"};
    if (h->ev[3] && cudaEventElapsedTime(&at, h->ev[3], h->comm_marks.ev[i]) == cudaSuccess)
      fprintf(stderr, "[hb timeline] rank %d comm %s at %.3f ms\n", h->rank, what[h->comm_marks.tag[i] - 100], at);
    else cudaGetLastError();
  }
  h->comm_marks.n = 0;
  h->timing.total_ms = h->timing.land_ms + h->timing.river_ms;
  h->pending = false;
  if (h->h_wave_err && *h->h_wave_err) {
    *h->h_wave_err = 0;
    cudaMemset(h->d_wave_ctl.p, 0, 2 * sizeof(int));
    return h->fail(HB_ECUDA, "routing wavefront: a task waited beyond the spin limit for its upstream reaches "
                             "(results of this window are invalid; HB_OPT_RIVER_SWEEP = HB_RIVER_LAUNCHES avoids the "
                             "persistent kernel)");
  }
  return HB_OK;
}

int hb_run_async(hb_engine* h, int flags) {
  if (!h) return HB_EINVAL;
  return enqueue_run(h, flags);
}

int hb_run(hb_engine* h, int flags) {
  if (!h) return HB_EINVAL;
  int rc = enqueue_run(h, flags);
  if (rc) return rc;
  return hb_wait(h);
}

__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1.0, a2 = a0 + 2.0, a3 = a0 + 3.0, a4 = a0 + 4.0, a5 = a0 + 5.0,
         a6 = a0 + 6.0, a7 = a0 + 7.0;
  const double m = 1.0000001, c = 1e-7;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (s == 123.456) out[0] = s;  // never true; keeps the chain alive
}

int hb_measure_fp64_peak(hb_engine* h, double* tflops) {
  if (!h || !tflops) return HB_EINVAL;
  HB_CUDA(h, cudaSetDevice(h->device));
  HB_CUDA(h, h->d_commbuf.reserve(1));
  const int iters = 1 << 15, block = 256, grid = h->prop.multiProcessorCount * 16;
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    HB_CUDA(h, cudaEventRecord(h->ev[0], h->stream));
    fp64_peak_kernel<<<grid, block, 0, h->stream>>>(h->d_commbuf.p, iters);
    HB_CUDA(h, cudaGetLastError());
    HB_CUDA(h, cudaEventRecord(h->ev[1], h->stream));
    HB_CUDA(h, cudaStreamSynchronize(h->stream));
    float ms = 0.f;
    HB_CUDA(h, cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
    const double tf = 2.0 * 8.0 * (double)iters * block * grid / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  *tflops = best;
  return HB_OK;
}

static int join_streams(hb_engine* h) {
  // everything enqueued so far on the routing and copy streams becomes a dependency of the land stream
  HB_CUDA(h, cudaEventRecord(h->ev_join[0], h->s_river));
  HB_CUDA(h, cudaEventRecord(h->ev_join[1], h->s_copy));
  HB_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_join[0], 0));
  HB_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_join[1], 0));
  HB_CUDA(h, cudaEventRecord(h->ev_join[0], h->s_comm));   // (the wait above already holds the first recording)
  HB_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_join[0], 0));
  for (auto& st : h->s_slab) {
    HB_CUDA(h, cudaEventRecord(h->ev_join[0], st));
    HB_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_join[0], 0));
  }
  HB_CUDA(h, cudaEventRecord(h->ev_join[0], h->s_join));
  HB_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_join[0], 0));
  return HB_OK;
}

int hb_timer_start(hb_engine* h) {
  if (!h) return HB_EINVAL;
  HB_CUDA(h, cudaSetDevice(h->device));
  int rc = join_streams(h);
  if (rc) return rc;
  HB_CUDA(h, cudaEventRecord(h->ev[3], h->stream));
  return HB_OK;
}

int hb_timer_stop(hb_engine* h, double* ms) {
  if (!h || !ms) return HB_EINVAL;
  HB_CUDA(h, cudaSetDevice(h->device));
  if (!h->ev_stop) HB_CUDA(h, cudaEventCreate(&h->ev_stop));
  int rc = join_streams(h);
  if (rc) return rc;
  HB_CUDA(h, cudaEventRecord(h->ev_stop, h->stream));
  HB_CUDA(h, cudaEventSynchronize(h->ev_stop));
  float f = 0.f;
  HB_CUDA(h, cudaEventElapsedTime(&f, h->ev[3], h->ev_stop));
  *ms = f;
  return HB_OK;
}

int hb_get_timing(hb_engine* h, hb_timing* out) {
  if (!h || !out) return HB_EINVAL;
  if (h->pending) { int rc = hb_wait(h); if (rc) return rc; }
  {
    unsigned long long c[HB_CNT_COUNT];
    HB_CUDA(h, cudaSetDevice(h->device));
    HB_CUDA(h, cudaMemcpyAsync(c, h->d_counters.p, sizeof c, cudaMemcpyDeviceToHost, h->stream));
    HB_CUDA(h, cudaMemsetAsync(h->d_counters.p, 0, sizeof c, h->stream));
    HB_CUDA(h, cudaStreamSynchronize(h->stream));
    h->timing.zone_warpsteps = (int64_t)c[HB_CNT_ZONE_WARPSTEPS];
    h->timing.zone_warpsteps_dry = (int64_t)c[HB_CNT_ZONE_WARPSTEPS_DRY];
    h->timing.zone_steps = (int64_t)c[HB_CNT_ZONE_STEPS];
    h->timing.zone_steps_wet = (int64_t)c[HB_CNT_ZONE_STEPS_WET];
    h->timing.elem_steps = (int64_t)c[HB_CNT_ELEM_STEPS];
    h->timing.elem_steps_run = (int64_t)c[HB_CNT_ELEM_STEPS_RUN];
    h->timing.substeps_pow = (int64_t)c[HB_CNT_SUBSTEPS_POW];
    h->timing.elem_warpsteps = (int64_t)c[HB_CNT_ELEM_WARPSTEPS];
    h->timing.elem_warpsteps_run = (int64_t)c[HB_CNT_ELEM_WARPSTEPS_RUN];
    h->timing.spec_reruns = (int64_t)c[HB_CNT_SPEC_RERUN];
  }
  *out = h->timing;
  // sums restart with the next run
  const int64_t n_levels = h->timing.n_levels, n_fast = h->timing.n_fast, n_fused = h->timing.n_fused,
                n_spec = h->timing.n_spec;
  h->timing = hb_timing{};
  h->timing.n_levels = n_levels;
  h->timing.n_fast = n_fast;
  h->timing.n_fused = n_fused;
  h->timing.n_spec = n_spec;
  return HB_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// results
// ---------------------------------------------------------------------------------------------------------------------
static int download(hb_engine* h, double* dst, const double* src, size_t n) {
  int rc = hb_wait(h);  // results of an asynchronous run must be complete
  if (rc) return rc;
  if (n) {
    HB_CUDA(h, cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    HB_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  return HB_OK;
}

int hb_get_land_outflow(hb_engine* h, double* rows) {
  if (!h) return HB_EINVAL;
  HB_REQUIRE(h, h->have_results, HB_ESTATE, "hb_get_land_outflow: no results (call hb_run)");
  HB_REQUIRE(h, rows || h->n_elem * h->nt == 0, HB_EINVAL, "hb_get_land_outflow: NULL argument");
  HB_CUDA(h, cudaSetDevice(h->device));
  return download(h, rows, h->d_land_rows[h->cur].p, (size_t)h->n_elem * h->nt);
}

__global__ void gather_rows_kernel(const double* __restrict__ src, double* __restrict__ dst,
                                   const int32_t* __restrict__ sel, int64_t n_sel, int64_t nt) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_sel * nt) return;
  const int64_t r = i / nt, t = i - r * nt;
  dst[i] = src[(int64_t)sel[r] * nt + t];
}

static int node_series(hb_engine* h, int64_t n_sel, const int32_t* nodes, double* rows, bool async) {
  HB_REQUIRE(h, h->have_river && h->have_routed, HB_ESTATE, "hb_get_node_series: no routing results (call hb_run)");
  HB_CUDA(h, cudaSetDevice(h->device));
  const int c = h->cur;
  const double* src = h->d_node_rows[c].p;
  size_t n = (size_t)h->n_node * h->nt;
  HB_CUDA(h, cudaStreamWaitEvent(h->s_copy, h->ev_river_done[c], 0));
  if (n_sel == 0) {
    HB_REQUIRE(h, rows || n == 0, HB_EINVAL, "hb_get_node_series: NULL argument");
  } else {
    HB_REQUIRE(h, n_sel > 0 && nodes && rows, HB_EINVAL, "hb_get_node_series: invalid selection");
    for (int64_t i = 0; i < n_sel; ++i)
      HB_REQUIRE(h, nodes[i] >= 0 && nodes[i] < h->n_node, HB_EINVAL, "hb_get_node_series: node %d out of range", nodes[i]);
    n = (size_t)n_sel * h->nt;
    HB_CUDA(h, h->d_sel.reserve((size_t)n_sel));
    HB_CUDA(h, h->d_gather[c].reserve(n));
    // the selection is small: a synchronous copy keeps the caller's array free to go
    HB_CUDA(h, cudaMemcpyAsync(h->d_sel.p, nodes, (size_t)n_sel * sizeof(int32_t), cudaMemcpyHostToDevice, h->s_copy));
    if (n) {
      gather_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->s_copy>>>(src, h->d_gather[c].p, h->d_sel.p, n_sel,
                                                                            h->nt);
      HB_CUDA(h, cudaGetLastError());
    }
    src = h->d_gather[c].p;
  }
  if (n) HB_CUDA(h, cudaMemcpyAsync(rows, src, n * sizeof(double), cudaMemcpyDeviceToHost, h->s_copy));
  HB_CUDA(h, cudaEventRecord(h->ev_copy_done[c], h->s_copy));
  if (!async) return hb_wait(h);
  h->pending = true;
  return HB_OK;
}

int hb_get_node_series(hb_engine* h, int64_t n_sel, const int32_t* nodes, double* rows) {
  if (!h) return HB_EINVAL;
  return node_series(h, n_sel, nodes, rows, false);
}

int hb_get_node_series_async(hb_engine* h, int64_t n_sel, const int32_t* nodes, double* rows) {
  if (!h) return HB_EINVAL;
  return node_series(h, n_sel, nodes, rows, true);
}

int hb_node_statistics(hb_engine* h, int64_t n_pair, const int32_t* nodes, const int32_t* obs_row, int64_t n_obs,
                       const double* obs, int skip_nan, double* out) {
  if (!h) return HB_EINVAL;
  HB_REQUIRE(h, h->have_river && h->have_routed, HB_ESTATE, "hb_node_statistics: no routing results (call hb_run)");
  HB_REQUIRE(h, n_pair >= 0 && n_obs >= 0 && (n_pair == 0 || (nodes && obs_row && obs && out)), HB_EINVAL,
             "hb_node_statistics: invalid argument");
  for (int64_t i = 0; i < n_pair; ++i)
    HB_REQUIRE(h, nodes[i] >= 0 && nodes[i] < h->n_node && obs_row[i] >= 0 && obs_row[i] < n_obs, HB_EINVAL,
               "hb_node_statistics: pair %lld refers to node %d / observation row %d", (long long)i, nodes[i], obs_row[i]);
  if (n_pair == 0) return HB_OK;
  HB_CUDA(h, cudaSetDevice(h->device));
  int rc = hb_wait(h);
  if (rc) return rc;
  if ((rc = upload(h, h->d_stat_nodes, nodes, (size_t)n_pair)) || (rc = upload(h, h->d_stat_rows, obs_row, (size_t)n_pair)) ||
      (rc = upload(h, h->d_obs, obs, (size_t)n_obs * h->nt)))
    return rc;
  HB_CUDA(h, h->d_stats.reserve((size_t)n_pair * HB_ST_COUNT));
  node_statistics_kernel<<<(unsigned)n_pair, 256, 0, h->stream>>>(h->d_node_rows[h->cur].p, h->d_obs.p, h->d_stat_nodes.p,
                                                                  h->d_stat_rows.p, h->nt, skip_nan, h->d_stats.p);
  HB_CUDA(h, cudaGetLastError());
  HB_CUDA(h, cudaMemcpyAsync(out, h->d_stats.p, (size_t)n_pair * HB_ST_COUNT * sizeof(double), cudaMemcpyDeviceToHost,
                             h->stream));
  HB_CUDA(h, cudaStreamSynchronize(h->stream));
  return HB_OK;
}

int hb_get_reach_outflow(hb_engine* h, double* rows) {
  if (!h) return HB_EINVAL;
  HB_REQUIRE(h, h->have_river && h->have_routed, HB_ESTATE, "hb_get_reach_outflow: no routing results");
  HB_CUDA(h, cudaSetDevice(h->device));
  return download(h, rows, h->d_reach_out[h->cur].p, (size_t)h->n_reach * h->nt);
}

int hb_get_reach_state_series(hb_engine* h, double* out) {
  if (!h) return HB_EINVAL;
  HB_REQUIRE(h, h->have_river && h->have_routed && h->river_state_series && h->d_dis_series[h->cur].p, HB_ESTATE,
             "hb_get_reach_state_series: not recorded (set HB_OPT_RIVER_STATE_SERIES before hb_run)");
  HB_CUDA(h, cudaSetDevice(h->device));
  return download(h, out, h->d_dis_series[h->cur].p, (size_t)h->nt * h->n_seg);
}

int hb_get_mct_series(hb_engine* h, int which, double* out) {
  if (!h) return HB_EINVAL;
  HB_REQUIRE(h, which >= 0 && which < HB_MS_COUNT, HB_EINVAL, "hb_get_mct_series: unknown sequence %d", which);
  HB_REQUIRE(h, h->have_river && h->have_routed, HB_ESTATE, "hb_get_mct_series: no routing results");
  const unsigned mask = h->mct_series_run[h->cur];
  HB_REQUIRE(h, (mask >> which & 1u) && h->d_mct_series[h->cur].p, HB_ESTATE,
             "hb_get_mct_series: sequence %d was not recorded (set its bit in HB_OPT_RIVER_MCT_SERIES before hb_run; "
             "the river needs musk_mct reaches)", which);
  HB_CUDA(h, cudaSetDevice(h->device));
  const size_t slot = (size_t)__builtin_popcount(mask & ((1u << which) - 1u)), n = (size_t)h->nt * h->n_seg;
  return download(h, out, h->d_mct_series[h->cur].p + slot * n, n);
}

int hb_get_reach_inflow(hb_engine* h, double* rows) {
  if (!h) return HB_EINVAL;
  HB_REQUIRE(h, h->have_river && h->have_routed, HB_ESTATE, "hb_get_reach_inflow: no routing results");
  HB_CUDA(h, cudaSetDevice(h->device));
  return download(h, rows, h->d_reach_in[h->cur].p, (size_t)h->n_reach * h->nt);
}

int hb_get_series(hb_engine* h, int series, double* out) {
  if (!h) return HB_EINVAL;
  HB_REQUIRE(h, series >= 0 && series < HB_S_COUNT, HB_EINVAL, "hb_get_series: unknown series %d", series);
  HB_REQUIRE(h, h->have_results && h->series_on[series] && h->d_series[series].p, HB_ESTATE,
             "hb_get_series: series %d was not selected before hb_run", series);
  HB_CUDA(h, cudaSetDevice(h->device));
  const int level = HB_S_LEVEL(series);
  const int64_t width = level == 0 ? h->n_zone : level == 1 ? h->n_snow : level == 2 ? h->n_elem : h->n_uh;
  return download(h, out, h->d_series[series].p, (size_t)h->nt * width);
}

// ---------------------------------------------------------------------------------------------------------------------
// multi-GPU boundary hand-over
// ---------------------------------------------------------------------------------------------------------------------
#define HB_NCCL(h, call)                                                                                   \
  do {                                                                                                     \
    int rc__ = (call);                                                                                     \
    if (rc__ != 0) return (h)->fail(HB_ENCCL, "%s failed: %s", #call, (h)->nccl.GetErrorString(rc__));     \
  } while (0)

int hb_comm_unique_id(hb_engine* h, void* id) {
  if (!h || !id) return HB_EINVAL;
  std::string err;
  if (!h->nccl.load(err)) return h->fail(HB_ENCCL, "%s", err.c_str());
  static_assert(sizeof(ncclUniqueId) == HB_NCCL_ID_BYTES, "ncclUniqueId size");
  HB_NCCL(h, h->nccl.GetUniqueId((ncclUniqueId*)id));
  return HB_OK;
}

int hb_comm_init(hb_engine* h, int rank, int world, const void* id) {
  if (!h || !id) return HB_EINVAL;
  HB_REQUIRE(h, world >= 1 && rank >= 0 && rank < world, HB_EINVAL, "hb_comm_init: invalid rank/world");
  std::string err;
  if (!h->nccl.load(err)) return h->fail(HB_ENCCL, "%s", err.c_str());
  HB_CUDA(h, cudaSetDevice(h->device));
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof uid);
  HB_NCCL(h, h->nccl.CommInitRank(&h->comm, world, uid, rank));
  h->rank = rank;
  h->world = world;
  return HB_OK;
}

__global__ void scatter_rows_kernel(const double* __restrict__ src, double* __restrict__ dst,
                                    const int32_t* __restrict__ sel, int64_t n_sel, int64_t nt) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_sel * nt) return;
  const int64_t r = i / nt, t = i - r * nt;
  dst[(int64_t)sel[r] * nt + t] = src[i];
}

// staging slot of the next send/receive call of the current window; the row list is uploaded only when it differs from what
// the slot holds (a sharded run repeats the same lists every window: no per-window host-to-device copy on the comm stream)
static int comm_slot(hb_engine* h, const int32_t* rows, size_t n_rows, size_t n_values, double** stage, int32_t** rows_dev) {
  const int c = h->cur;
  const size_t slot = (size_t)h->comm_calls[c]++;
  if (h->d_comm_stage[c].size() <= slot) {
    h->d_comm_stage[c].resize(slot + 1); h->d_comm_rows[c].resize(slot + 1); h->comm_rows_host[c].resize(slot + 1);
  }
  HB_CUDA(h, h->d_comm_stage[c][slot].reserve(n_values));
  HB_CUDA(h, h->d_comm_rows[c][slot].reserve(n_rows));
  std::vector<int32_t>& held = h->comm_rows_host[c][slot];
  if (held.size() != n_rows || !std::equal(held.begin(), held.end(), rows)) {
    held.assign(rows, rows + n_rows);
    // (the previous message of this slot — two windows ago — must be through before its row list changes)
    HB_CUDA(h, cudaStreamSynchronize(h->s_comm));
    HB_CUDA(h, cudaMemcpy(h->d_comm_rows[c][slot].p, held.data(), n_rows * sizeof(int32_t), cudaMemcpyHostToDevice));
  }
  *stage = h->d_comm_stage[c][slot].p;
  *rows_dev = h->d_comm_rows[c][slot].p;
  return HB_OK;
}

int hb_comm_send_rows(hb_engine* h, int peer, int kind, int64_t n, const int32_t* rows) {
  if (!h) return HB_EINVAL;
  HB_REQUIRE(h, h->comm, HB_ESTATE, "hb_comm_send_rows: call hb_comm_init first");
  HB_REQUIRE(h, h->have_river && h->have_routed, HB_ESTATE, "hb_comm_send_rows: no routing results");
  HB_REQUIRE(h, peer >= 0 && peer < h->world && peer != h->rank && n >= 0 && (n == 0 || rows) &&
                    (kind == HB_ROW_NODE || kind == HB_ROW_REACH),
             HB_EINVAL, "hb_comm_send_rows: invalid argument");
  HB_CUDA(h, cudaSetDevice(h->device));
  const int64_t limit = kind == HB_ROW_NODE ? h->n_node : h->n_reach;
  const double* base = kind == HB_ROW_NODE ? h->d_node_rows[h->cur].p : h->d_reach_out[h->cur].p;
  for (int64_t i = 0; i < n; ++i)
    HB_REQUIRE(h, rows[i] >= 0 && rows[i] < limit, HB_EINVAL, "hb_comm_send_rows: row %d out of range", rows[i]);
  if (n == 0 || h->nt == 0) return HB_OK;
  // all NCCL traffic of this engine runs on ONE stream (s_comm); the send waits for the routing of the window
  HB_CUDA(h, cudaStreamWaitEvent(h->s_comm, h->ev_river_done[h->cur], 0));
  static const bool timeline = getenv("HB_DEBUG_TIMELINE") != nullptr;
  if (timeline) h->comm_marks.add(100, h->s_comm);     // the routing of the window is through
  const double* src = base + (int64_t)rows[0] * h->nt;
  if (n > 1) {
    // the rows of a peer travel as ONE message: packed by a gather kernel behind the routing (a basin cut into sub-basins
    // hands over hundreds of short series per window; one NCCL operation per row would cost more than the transfer)
    double* stage = nullptr;
    int32_t* rows_dev = nullptr;
    int rc = comm_slot(h, rows, (size_t)n, (size_t)n * h->nt, &stage, &rows_dev);
    if (rc) return rc;
    const size_t total = (size_t)n * h->nt;
    gather_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, h->s_comm>>>(base, stage, rows_dev, n, h->nt);
    HB_CUDA(h, cudaGetLastError());
    src = stage;
  }
  if (timeline) h->comm_marks.add(101, h->s_comm);     // rows packed
  HB_NCCL(h, h->nccl.Send(src, (size_t)n * h->nt, kNcclFloat64, peer, h->comm, h->s_comm));
  if (timeline) h->comm_marks.add(102, h->s_comm);     // sent
  HB_CUDA(h, cudaEventRecord(h->ev_send_done[h->cur], h->s_comm));
  h->pending = true;
  return HB_OK;
}

int hb_comm_recv_ext(hb_engine* h, int peer, int64_t n, const int32_t* ext_rows) {
  if (!h) return HB_EINVAL;
  HB_REQUIRE(h, h->comm, HB_ESTATE, "hb_comm_recv_ext: call hb_comm_init first");
  HB_REQUIRE(h, h->have_river && h->have_forcing, HB_ESTATE, "hb_comm_recv_ext: river/window not set");
  HB_REQUIRE(h, peer >= 0 && peer < h->world && peer != h->rank && n >= 0 && (n == 0 || ext_rows), HB_EINVAL,
             "hb_comm_recv_ext: invalid argument");
  HB_CUDA(h, cudaSetDevice(h->device));
  bool contiguous = true;
  for (int64_t i = 0; i < n; ++i) {   // (before any NCCL call: no error path may leave an operation half posted)
    HB_REQUIRE(h, ext_rows[i] >= 0 && ext_rows[i] < h->n_ext, HB_EINVAL, "hb_comm_recv_ext: row %d out of range", ext_rows[i]);
    if (i > 0 && ext_rows[i] != ext_rows[i - 1] + 1) contiguous = false;
  }
  HB_CUDA(h, h->d_ext_rows[h->cur].reserve((size_t)h->n_ext * h->nt));
  h->ext_async[h->cur] = true;
  h->have_ext = true;
  if (n == 0 || h->nt == 0) {
    HB_CUDA(h, cudaEventRecord(h->ev_ext_ready[h->cur], h->s_comm));
    return HB_OK;
  }
  // the routing of two windows ago was the last reader of this buffer set
  HB_CUDA(h, cudaStreamWaitEvent(h->s_comm, h->ev_river_done[h->cur], 0));
  // an NCCL receive occupies SMs from its launch until the peer's data is in: it is posted behind the runoff generation of
  // the window (when there is one in flight), i.e. about when the upstream ranks start routing the same window
  if (h->have_results) HB_CUDA(h, cudaStreamWaitEvent(h->s_comm, h->ev_land_done[h->cur], 0));
  // ONE message per peer (see hb_comm_send_rows).  Consecutive external rows — what hydpy_b200.shard hands out — are
  // received straight into the buffer the reach kernels read; any other selection goes through a staging buffer
  if (contiguous) {
    HB_NCCL(h, h->nccl.Recv(h->d_ext_rows[h->cur].p + (int64_t)ext_rows[0] * h->nt, (size_t)n * h->nt, kNcclFloat64, peer,
                            h->comm, h->s_comm));
  } else {
    double* stage = nullptr;
    int32_t* rows_dev = nullptr;
    int rc = comm_slot(h, ext_rows, (size_t)n, (size_t)n * h->nt, &stage, &rows_dev);
    if (rc) return rc;
    HB_NCCL(h, h->nccl.Recv(stage, (size_t)n * h->nt, kNcclFloat64, peer, h->comm, h->s_comm));
    const size_t total = (size_t)n * h->nt;
    scatter_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, h->s_comm>>>(stage, h->d_ext_rows[h->cur].p, rows_dev, n,
                                                                               h->nt);
    HB_CUDA(h, cudaGetLastError());
  }
  {
    static const bool timeline = getenv("HB_DEBUG_TIMELINE") != nullptr;
    if (timeline) h->comm_marks.add(103, h->s_comm);   // received
  }
  HB_CUDA(h, cudaEventRecord(h->ev_ext_ready[h->cur], h->s_comm));
  h->pending = true;
  return HB_OK;
}

int hb_comm_barrier(hb_engine* h) {
  if (!h) return HB_EINVAL;
  HB_REQUIRE(h, h->comm, HB_ESTATE, "hb_comm_barrier: call hb_comm_init first");
  HB_CUDA(h, cudaSetDevice(h->device));
  HB_CUDA(h, h->d_commbuf.reserve(1));
  HB_CUDA(h, cudaMemsetAsync(h->d_commbuf.p, 0, sizeof(double), h->s_comm));
  HB_NCCL(h, h->nccl.AllReduce(h->d_commbuf.p, h->d_commbuf.p, 1, kNcclFloat64, kNcclSum, h->comm, h->s_comm));
  HB_CUDA(h, cudaStreamSynchronize(h->s_comm));
  return HB_OK;
}

}  // extern "C"
