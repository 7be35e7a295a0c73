// ctx.cu -- context, wave planning and the C ABI entry points of libfrieda_b200.so.
//
// Call sequences mirror the reference glue:
//   frieda_commit*      : src/commit.rs:11-23   pack -> LDE -> Merkle root
//   frieda_fri_commit*  : src/proof.rs:38-57    + channel, FriProver::commit
//   frieda_prove*       : src/proof.rs:32-77    + grind, mix_u64, decommit, evaluations
// A batched call is cut into waves that fit the device workspace; within a wave every kernel is
// launched once over all blobs (grid.y / grid.z = blob).  There is no CPU fallback: all stages run
// as CUDA kernels on the context's stream.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <atomic>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include "../../include/frieda_b200.h"
#include "ctx_internal.h"
#include "host_math.hpp"
#include "kernels.cuh"
#include "split_plan.hpp"

using namespace frieda;

namespace {

thread_local std::string g_create_error;

struct Geom {
  size_t len = 0;
  uint32_t n_felts = 0, p = 0, beta = 0, D = 0;
  // proof shape
  uint32_t log_last = 0, last_log = 0, n_inner = 0, n_layers = 0;
};

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// FRIEDA_TRACE=1: host-side stage timings of the proof path on stderr (development aid)
struct StageTrace {
  bool on;
  std::chrono::steady_clock::time_point t0;
  StageTrace() : on(std::getenv("FRIEDA_TRACE") != nullptr), t0(std::chrono::steady_clock::now()) {}
  void mark(const char *what) {
    if (!on) return;
    auto t1 = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[frieda trace] %-28s %8.1f us\n", what,
                 std::chrono::duration<double, std::micro>(t1 - t0).count());
    t0 = t1;
  }
};

struct Bump {
  size_t off = 0;
  size_t take(size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  }
};

struct Plan {
  Geom g;
  size_t B = 0;
  uint32_t levels_cfg = 10;  // Merkle per-pass depth chosen for this call
  uint32_t nq = 0;  // n_queries (prove)
  bool keep = false, prove = false, fri = false;
  size_t in_stride = 0;  // bytes between staged blobs
  size_t o_in = 0, o_in2 = 0, o_coef = 0, o_chan = 0, o_alpha = 0, o_roots = 0, o_last = 0, o_err = 0, o_seeds = 0;
  size_t o_cols[34] = {0}, cols_stride[34] = {0};  // u32 elements per blob
  size_t o_tree[34] = {0}, tree_stride[34] = {0};  // 32-byte slots per blob
  size_t o_best = 0, o_next = 0, o_queries = 0, o_nuniq = 0, o_counts = 0, o_lvl = 0, o_offsets = 0, o_totals = 0,
         o_evals = 0;
  size_t total = 0;
};

uint32_t layer_log(const Geom &g, uint32_t layer) { return g.D - layer; }

// Levels reduced inside one CTA of the Merkle bottom/middle passes.  Deep (10: 1024 leaves -> 1
// node) minimises launches but parks most warps at barriers during the last 8 levels; shallow
// (2: 1024 -> 256) keeps every thread hashing in every level at the cost of more passes.
// levels_cfg: bits 0-7 = levels per pass.  LV_LATENCY (a few blobs have the GPU to themselves: the tree is a chain of
// dependent compressions, not a throughput problem) selects the latency form of the Merkle passes (merkle.cu) and
// chunks small enough that the bottom pass leaves about 2^t nodes per blob (bits 16-23 = t), i.e. roughly one CTA
// per SM over the blobs of the call.
constexpr uint32_t LV_LATENCY = 0x100;
uint32_t lv_levels(uint32_t levels_cfg) { return levels_cfg & 0xffu; }
uint32_t chunk_log_for(uint32_t d, uint32_t levels_cfg) {
  uint32_t chunk = d < 10 ? d : 10;
  if (levels_cfg & LV_LATENCY) {
    const uint32_t t = (levels_cfg >> 16) & 0xffu;
    uint32_t c = d > t ? d - t : 0;
    if (c < 7) c = 7;
    if (c < chunk) chunk = c;
  }
  return chunk;
}
uint32_t pass_levels(uint32_t d, uint32_t levels_cfg) {
  const uint32_t chunk = chunk_log_for(d, levels_cfg), lv = lv_levels(levels_cfg);
  return lv < chunk ? lv : chunk;
}
size_t tree_slots(uint32_t d, bool keep, uint32_t levels_cfg) {
  // kept trees hold every level; truncated trees only the levels above the bottom pass
  // (tail layers: just the root in slot 1)
  if (keep) return (size_t)2 << d;
  return (size_t)2 << (d - pass_levels(d, levels_cfg));
}

}  // namespace

struct frieda_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  // host-buffer entry points stage the next wave's input on copy_stream while the current wave
  // computes: ev_copied[b] = staging buffer b holds its wave, ev_free[b] = its pack kernel has read it
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
  std::string err;
  uint64_t launches = 0;
  size_t ws_limit = 0;
  uint32_t merkle_levels_big = 3;    // per-pass depth for large batches (env FRIEDA_MERKLE_LEVELS)
  uint32_t merkle_levels_small = 10; // per-pass depth when the grid would not fill the GPU anyway
  uint32_t merkle_latency_max_chunk = 8;  // (env FRIEDA_MERKLE_LATENCY_CHUNK, for probing)
  bool merkle_latency_form = true;   // latency form of the Merkle passes for those calls (env FRIEDA_MERKLE_LATENCY=0: off)
  uint32_t levels_for(size_t n_blobs, uint32_t d) const {
    // CTAs of the bottom pass; below ~4 waves of 148 SMs x 4 CTAs the launch count matters more
    size_t ctas = n_blobs << (d > 10 ? d - 10 : 0);
    if (ctas >= 2368) return merkle_levels_big;
    if (merkle_levels_small != 10 || !merkle_latency_form) return merkle_levels_small;
    uint32_t lg = 0;  // floor(log2(n_blobs))
    while ((n_blobs >> (lg + 1)) != 0) lg++;
    const uint32_t t = lg >= 7 ? 0 : 7 - lg;
    return merkle_levels_small | 0x100u /* LV_LATENCY */ | (t << 16);
  }
  // twiddle cache
  bool tw_valid = false;
  uint32_t tw_K = 0;
  uint32_t *d_tw = nullptr, *d_itw = nullptr, *d_tw2 = nullptr;
  GenPowers gp;
  // workspace arena
  uint8_t *arena = nullptr;
  size_t arena_bytes = 0;
  // small result staging
  uint8_t *d_scratch = nullptr;  // 8 KiB: [0, 4 KiB) top-of-tree scratch (2 x 64 slots), then small flags
  // "invalid degree" flag of the asynchronous *_device entry points: set by the tail kernel, read and cleared
  // by frieda_ctx_take_error (the host-buffer entry points use a flag inside the wave's workspace instead)
  int *d_async_err() const { return reinterpret_cast<int *>(d_scratch + 4096 + 64); }
  // grow-only buffers of the proof path: gathered witnesses on the device, pinned readback on the host
  // (both double-buffered: wave i is read back on the copy stream and assembled while wave i+1 computes)
  uint8_t *d_gather[2] = {nullptr, nullptr};
  size_t d_gather_bytes[2] = {0, 0};
  uint8_t *h_pinned[2] = {nullptr, nullptr};
  size_t h_pinned_bytes[2] = {0, 0};
  cudaEvent_t ev_gathered[2] = {nullptr, nullptr};  // wave's witnesses are in d_gather[b], small results read back
  cudaEvent_t ev_readback[2] = {nullptr, nullptr};  // d_gather[b] has reached h_pinned[b]
  // per-kernel timing with CUDA events on the launching stream (bench.py's roofline)
  bool profiling = false;
  struct ProfRec {
    const char *name;
    cudaEvent_t a, b;
  };
  std::vector<ProfRec> prof_recs;
  std::vector<cudaEvent_t> prof_pool;
  void prof_begin(const char *name) {
    if (!profiling) return;
    ProfRec r{name, nullptr, nullptr};
    for (cudaEvent_t *e : {&r.a, &r.b}) {
      if (!prof_pool.empty()) {
        *e = prof_pool.back();
        prof_pool.pop_back();
      } else {
        cudaEventCreate(e);
      }
    }
    cudaEventRecord(r.a, stream);
    prof_recs.push_back(r);
  }
  void prof_end() {
    if (!profiling || prof_recs.empty()) return;
    cudaEventRecord(prof_recs.back().b, stream);
  }
  // FRI commit phase of ONE blob split over `world` ranks (frieda_fri_split_*): geometry and workspace offsets
  struct SplitFri {
    bool active = false;
    Geom g;
    uint32_t rank = 0, world = 1, gl = 0;
    uint32_t n_split = 0;        // layers 0 .. n_split-1 are committed rank-locally + one exchange of subtree roots
    uint32_t next_layer = 0;     // next layer expected by frieda_fri_split_layer
    bool handed_off = false;
    bool keep = false;           // trees kept whole: a proof follows (frieda_fri_split_decommit)
    bool finished = false;       // FriProver::commit is complete, the state is still resident
    bool peer_timeout_check = false;  // the layers ran over peer memory: `finish` reads the barrier-timeout flag
    frieda_pcs_config cfg{};
    size_t o_toptree = 0;        // keep: per split layer the top log2(world) levels (2 * world slots each)
    size_t o_best = 0, o_next = 0, o_queries = 0, o_nuniq = 0;
    Plan rem;                    // the unsplit remainder as a one-blob wave (valid once finished)
    size_t o_coef = 0, o_chan = 0, o_alpha = 0, o_roots = 0, o_last = 0, o_err = 0, o_seed = 0, o_top = 0, o_sub = 0;
    size_t o_cols[34] = {0};     // local columns of layers 0 .. n_split-1 (4 x 2^(D - l - gl) u32)
    size_t o_tree[34] = {0}, tree_slots[34] = {0};
    size_t o_full = 0;           // workspace of the unsplit remainder (a Plan laid out behind the split state)
    uint32_t levels_cfg[34] = {0};
  } split;
  // introspection
  bool debug_keep = false;
  bool have_last = false;
  Plan last;

  int fail(cudaError_t e, const char *what, int line) {
    char buf[512];
    std::snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s [ctx.cu:%d]", (int)e, cudaGetErrorString(e), what, line);
    err = buf;
    cudaGetLastError();
    return FRIEDA_ERR_CUDA;
  }
  int fail_arg(const char *msg, int code = FRIEDA_ERR_ARG) {
    err = msg;
    return code;
  }
};

#define CU(call)                                                            \
  do {                                                                      \
    cudaError_t e_ = (call);                                                \
    if (e_ != cudaSuccess) return ctx->fail(e_, #call, __LINE__);           \
  } while (0)
#define KL(name, call, n)                                                   \
  do {                                                                      \
    ctx->prof_begin(name);                                                  \
    cudaError_t e_ = (call);                                                \
    ctx->prof_end();                                                        \
    if (e_ != cudaSuccess) return ctx->fail(e_, #call, __LINE__);           \
    ctx->launches += (n);                                                   \
  } while (0)

namespace {

int make_geom(frieda_ctx *ctx, size_t len, uint32_t log_blowup, Geom &g) {
  if (len >= ((size_t)1 << 31)) return ctx->fail_arg("input longer than 2^31 bytes is not supported");
  g.len = len;
  g.n_felts = (uint32_t)((len * 8 + 29) / 30);
  uint32_t lg = g.n_felts ? host::ceil_log2(g.n_felts) : 0;  // src/utils.rs:23 (f64 log2 of 0 casts to 0)
  if (lg < 2) lg = 2;
  g.p = lg - 2;
  g.beta = log_blowup;
  uint64_t D = (uint64_t)g.p + log_blowup;
  if (D == 0) return ctx->fail_arg("reference panics: Coset::half_odds(poly_log + log_blowup - 1) underflows",
                                   FRIEDA_ERR_PANIC);
  if (D > 28) return ctx->fail_arg("domain larger than 2^28 is not supported");
  g.D = (uint32_t)D;
  return FRIEDA_OK;
}

int make_geom_fri(frieda_ctx *ctx, size_t len, const frieda_pcs_config *cfg, Geom &g) {
  int rc = make_geom(ctx, len, cfg->log_blowup_factor, g);
  if (rc) return rc;
  g.log_last = cfg->log_last_layer_degree_bound;
  // commit_last_layer: assert_eq!(evaluation.len(), config.last_layer_domain_size())
  if ((uint64_t)g.p < 1 + (uint64_t)g.log_last)
    return ctx->fail_arg("reference panics: polynomial too small for log_last_layer_degree_bound", FRIEDA_ERR_PANIC);
  g.last_log = g.log_last + g.beta;
  g.n_inner = g.p - 1 - g.log_last;
  g.n_layers = g.n_inner + 1;
  if (g.D < 3) return ctx->fail_arg("FRI over a domain smaller than 8 points is not supported");
  if (g.last_log > TAIL_LAST_MAX)
    return ctx->fail_arg("log_last_layer_degree_bound + log_blowup_factor > 9 is not supported");
  if (g.n_layers > 30) return ctx->fail_arg("too many FRI layers");
  return FRIEDA_OK;
}

int ensure_twiddles(frieda_ctx *ctx, uint32_t K) {
  if (ctx->tw_valid && ctx->tw_K >= K) return FRIEDA_OK;
  if (ctx->d_tw) {
    CU(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->d_tw);
    cudaFree(ctx->d_itw);
    cudaFree(ctx->d_tw2);
    ctx->d_tw = ctx->d_itw = ctx->d_tw2 = nullptr;
    ctx->tw_valid = false;
  }
  size_t bytes = sizeof(uint32_t) << K;
  CU(cudaMalloc(&ctx->d_tw, bytes));
  CU(cudaMalloc(&ctx->d_itw, bytes));
  CU(cudaMalloc(&ctx->d_tw2, bytes));
  KL("twiddles", launch_twiddles(ctx->stream, ctx->gp, K, ctx->d_tw, ctx->d_itw, ctx->d_tw2), 1);
  ctx->tw_K = K;
  ctx->tw_valid = true;
  return FRIEDA_OK;
}

TwiddleTable table(const frieda_ctx *ctx) { return TwiddleTable{ctx->d_tw, ctx->d_itw, ctx->d_tw2, ctx->tw_K}; }

int ensure_arena(frieda_ctx *ctx, size_t bytes) {
  if (ctx->arena_bytes >= bytes) return FRIEDA_OK;
  if (ctx->arena) {
    CU(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->arena);
    ctx->arena = nullptr;
    ctx->arena_bytes = 0;
    ctx->have_last = false;
  }
  cudaError_t e = cudaMalloc(&ctx->arena, bytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    char buf[160];
    std::snprintf(buf, sizeof buf, "device workspace allocation of %zu bytes failed", bytes);
    ctx->err = buf;
    return FRIEDA_ERR_ALLOC;
  }
  ctx->arena_bytes = bytes;
  return FRIEDA_OK;
}

size_t workspace_budget(frieda_ctx *ctx) {
  if (ctx->ws_limit) return ctx->ws_limit;
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) {
    cudaGetLastError();
    return (size_t)8 << 30;
  }
  return (size_t)((double)(free_b + ctx->arena_bytes) * 0.8);
}

// Lays out one wave of B blobs.  stage_input: reserve a device copy of the input bytes.
void layout(Plan &pl, size_t B, bool stage_input) {
  const Geom &g = pl.g;
  pl.B = B;
  Bump bp;
  pl.in_stride = align_up(g.len ? g.len : 1, 16);
  if (stage_input) {
    pl.o_in = bp.take(B * pl.in_stride);
    pl.o_in2 = bp.take(B * pl.in_stride);  // double buffer: the next wave is uploaded during compute
  }
  pl.o_coef = bp.take(B * ((size_t)16 << g.p));
  const uint32_t n_cols_layers = pl.fri ? g.n_layers + 1 : 1;
  for (uint32_t l = 0; l < n_cols_layers; l++) {
    uint32_t d = (!pl.fri || l < g.n_layers) ? layer_log(g, l) : g.last_log;
    pl.cols_stride[l] = (size_t)4 << d;
    pl.o_cols[l] = bp.take(B * pl.cols_stride[l] * 4);
  }
  const uint32_t n_trees = pl.fri ? g.n_layers : 1;
  for (uint32_t l = 0; l < n_trees; l++) {
    pl.tree_stride[l] = tree_slots(layer_log(g, l), pl.keep, pl.levels_cfg);
    pl.o_tree[l] = bp.take(B * pl.tree_stride[l] * 32);
  }
  pl.o_roots = bp.take(B * (size_t)(pl.fri ? g.n_layers : 1) * 32);
  if (pl.fri) {
    pl.o_chan = bp.take(B * sizeof(Channel));
    pl.o_alpha = bp.take(B * g.n_layers * sizeof(QM31));
    pl.o_last = bp.take(B * (sizeof(QM31) << g.log_last));
    pl.o_err = bp.take(256);
    pl.o_seeds = bp.take(B * 8);
  }
  if (pl.prove) {
    pl.o_best = bp.take(B * 8);
    pl.o_next = bp.take(B * 8);
    pl.o_totals = bp.take(256);
  }
  pl.total = bp.off;
}

// per-wave extras of the prove path that depend on n_queries
void layout_prove_tail(Plan &pl, uint32_t n_queries) {
  Bump bp;
  bp.off = pl.total;
  pl.nq = n_queries;
  pl.o_queries = bp.take(pl.B * n_queries * 4);
  pl.o_nuniq = bp.take(pl.B * 4);
  pl.o_counts = bp.take(pl.B * pl.g.n_layers * 2 * 4);
  pl.o_offsets = bp.take(pl.B * pl.g.n_layers * 2 * 8);
  pl.o_lvl = bp.take(pl.B * pl.g.n_layers * (size_t)pl.g.D * 4);
  pl.o_evals = bp.take(pl.B * n_queries * sizeof(QM31));
  pl.total = bp.off;
}

size_t pick_wave(frieda_ctx *ctx, Plan &pl, size_t n, bool stage_input, uint32_t n_queries) {
  auto lay = [&](size_t B) {
    layout(pl, B, stage_input);
    if (pl.prove) layout_prove_tail(pl, n_queries);
  };
  size_t B = n > 32768 ? 32768 : n;
  // host input: at least 4 waves (of >= 128 blobs) so that all but the first upload hide behind compute
  // (8 waves measure the same end-to-end rate, 16 are 1.5 % slower)
  if (stage_input && n >= 512 && B > (n + 3) / 4) B = (n + 3) / 4;
  lay(B);
  // fast path: the workspace we already hold fits the whole call (cudaMemGetInfo costs milliseconds)
  if (pl.total <= ctx->arena_bytes && (!ctx->ws_limit || pl.total <= ctx->ws_limit)) return B;
  size_t budget = workspace_budget(ctx);
  if (pl.total <= budget) return B;
  lay(1);
  size_t per_blob = pl.total + 4096;
  size_t fit = budget / per_blob;
  if (fit < 1) fit = 1;
  if (fit < B) B = fit;
  lay(B);
  return B;
}

template <class T>
T *at(frieda_ctx *ctx, size_t off) {
  return reinterpret_cast<T *>(ctx->arena + off);
}

// Bottom pass (leaves from columns / fused fold), middle passes over existing nodes, then one CTA
// per blob for the top of the tree (+ the channel step).  `mp` carries the source/destination/tree
// pointers; sizes and depths are filled in here.
int run_tree(frieda_ctx *ctx, int src, MerkleBottomParams mp, uint32_t d, uint32_t levels_cfg, bool keep,
             size_t n_blobs, uint8_t *roots, size_t roots_stride, Channel *chan, QM31 *alpha, size_t alpha_stride) {
  mp.log = d;
  mp.chunk_log = chunk_log_for(d, levels_cfg);
  mp.levels = pass_levels(d, levels_cfg);
  mp.src_level = d;
  mp.write_all = keep ? 1 : 0;
  // the 128-thread latency form pays off for chunks of <= 256 leaves (one or two leaves per thread); larger chunks
  // keep the 256-thread form, whose second warp per scheduler hides the latency of the column loads
  mp.latency = ((levels_cfg & LV_LATENCY) && mp.chunk_log <= ctx->merkle_latency_max_chunk) ? 1 : 0;
  KL(src == SRC_COLS ? "merkle_bottom_cols" : src == SRC_FOLD_CIRCLE ? "fold_circle+merkle_bottom" : "fold_line+merkle_bottom",
     launch_merkle_bottom(ctx->stream, src, mp, n_blobs), 1);
  uint32_t u = d - mp.levels;
  while (u > 10) {
    MerkleBottomParams mm = mp;
    mm.src_cols = nullptr;
    mm.dst_cols = nullptr;
    mm.alpha = nullptr;
    mm.log = u;
    mm.src_level = u;
    mm.chunk_log = 10;
    mm.latency = 0;
    // shallow passes keep every thread hashing, but only pay off while the pass still fills the GPU: with fewer than
    // ~4 waves of CTAs (one big tree: config 5) a pass is latency-bound and fewer, deeper passes are faster
    const size_t ctas = n_blobs << (u - 10);
    mm.levels = (lv_levels(levels_cfg) < 10 && ctas >= 2368) ? lv_levels(levels_cfg) : 10;
    if (mm.levels > u - 10) mm.levels = u - 10;  // stop exactly where the top kernel takes over
    KL("merkle_mid", launch_merkle_bottom(ctx->stream, SRC_NODES, mm, n_blobs), 1);
    u -= mm.levels;
  }
  KL("merkle_top", launch_merkle_top(ctx->stream, mp.tree, mp.tree_stride, u, keep ? 1 : 0, roots, roots_stride, chan,
                                     alpha, alpha_stride, n_blobs),
     1);
  return FRIEDA_OK;
}

// Merkle tree over layer `layer` of a wave: existing columns (layer 0) or the fused fold of the
// previous layer.
int commit_tree(frieda_ctx *ctx, const Plan &pl, uint32_t layer, int src, Channel *chan, uint8_t *roots,
                size_t roots_stride) {
  const Geom &g = pl.g;
  const uint32_t d = layer_log(g, layer);
  MerkleBottomParams mp;
  std::memset(&mp, 0, sizeof mp);
  TwiddleTable tt = table(ctx);
  if (src == SRC_COLS) {
    mp.src_cols = at<uint32_t>(ctx, pl.o_cols[layer]);
    mp.src_stride = pl.cols_stride[layer];
  } else {
    mp.src_cols = at<uint32_t>(ctx, pl.o_cols[layer - 1]);
    mp.src_stride = pl.cols_stride[layer - 1];
    mp.dst_cols = at<uint32_t>(ctx, pl.o_cols[layer]);
    mp.dst_stride = pl.cols_stride[layer];
    mp.alpha = at<QM31>(ctx, pl.o_alpha) + (layer - 1);
    mp.alpha_stride = g.n_layers;
    // circle fold of the log-D evaluation: pairs (x, y) of the block of length 2^(D-2);
    // line fold of a log-(d+1) layer: block of length 2^d
    mp.itw_blk = src == SRC_FOLD_CIRCLE ? tt.iblk(1u << (g.D - 2)) : tt.iblk(1u << d);
  }
  mp.tree = at<uint8_t>(ctx, pl.o_tree[layer]);
  mp.tree_stride = pl.tree_stride[layer];
  QM31 *alpha = chan ? at<QM31>(ctx, pl.o_alpha) + layer : nullptr;
  return run_tree(ctx, src, mp, d, pl.levels_cfg, pl.keep, pl.B, roots, roots_stride, chan, alpha, g.n_layers);
}

CPoint half_initial_point(const Geom &g) { return host::point_from_index(half_odds_index(g.D - 1, 0)); }

// pack + LDE of the wave's blobs (device pointer d_in, stride bytes) into cols[0]
int lde_wave(frieda_ctx *ctx, const Plan &pl, const uint8_t *d_in, size_t stride, int staged_buf = -1) {
  const Geom &g = pl.g;
  uint32_t *coef = at<uint32_t>(ctx, pl.o_coef);
  KL("pack", launch_pack(ctx->stream, d_in, g.len, stride, pl.B, g.n_felts, g.p, coef), 1);
  if (staged_buf >= 0) CU(cudaEventRecord(ctx->ev_free[staged_buf], ctx->stream));  // staging buffer is reusable
  KL("lde", launch_lde(ctx->stream, coef, at<uint32_t>(ctx, pl.o_cols[0]), g.p, g.beta, pl.B, g.n_felts, table(ctx),
                half_initial_point(g)),
     (g.p > 15 ? 2 : 1));
  return FRIEDA_OK;
}

// Enqueues the upload of one wave (nb blobs) into staging buffer `buf` on the copy stream.
int stage_upload(frieda_ctx *ctx, const Plan &pl, int buf, const uint8_t *h_blobs, size_t stride, size_t nb) {
  uint8_t *dst = at<uint8_t>(ctx, buf ? pl.o_in2 : pl.o_in);
  const Geom &g = pl.g;
  CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_free[buf], 0));
  if (g.len) {
    if (stride == g.len && pl.in_stride == g.len) {
      CU(cudaMemcpyAsync(dst, h_blobs, nb * g.len, cudaMemcpyHostToDevice, ctx->copy_stream));
    } else {
      CU(cudaMemcpy2DAsync(dst, pl.in_stride, h_blobs, stride, g.len, nb, cudaMemcpyHostToDevice, ctx->copy_stream));
    }
  }
  CU(cudaEventRecord(ctx->ev_copied[buf], ctx->copy_stream));
  return FRIEDA_OK;
}
// Makes the compute stream wait for staging buffer `buf`; returns its device pointer.
int stage_acquire(frieda_ctx *ctx, const Plan &pl, int buf, const uint8_t **d_in, size_t *d_stride) {
  CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_copied[buf], 0));
  *d_in = at<uint8_t>(ctx, buf ? pl.o_in2 : pl.o_in);
  *d_stride = pl.in_stride;
  return FRIEDA_OK;
}
// Single-wave convenience used by the proof path.
int stage_in(frieda_ctx *ctx, const Plan &pl, const uint8_t *h_blobs, size_t stride, const uint8_t **d_in,
             size_t *d_stride) {
  // the previous user of buffer 0 is ordered on ctx->stream; make the copy stream follow it
  CU(cudaEventRecord(ctx->ev_free[0], ctx->stream));
  int rc = stage_upload(ctx, pl, 0, h_blobs, stride, pl.B);
  if (rc) return rc;
  return stage_acquire(ctx, pl, 0, d_in, d_stride);
}

// ---- commit ----------------------------------------------------------------------------
int commit_impl(frieda_ctx *ctx, const uint8_t *blobs, size_t len, size_t stride, size_t n, uint32_t log_blowup,
                uint8_t *roots_out, bool device_io) {
  if (!ctx) return FRIEDA_ERR_ARG;
  if ((!blobs && len) || !roots_out) return ctx->fail_arg("null pointer");
  if (n == 0) return FRIEDA_OK;
  if (n > 1 && stride < len) return ctx->fail_arg("blob_stride smaller than blob_len");
  CU(cudaSetDevice(ctx->device));
  Plan pl;
  int rc = make_geom(ctx, len, log_blowup, pl.g);
  if (rc) return rc;
  pl.keep = false;
  pl.levels_cfg = ctx->levels_for(n, pl.g.D);
  if (pl.g.D >= 3 && (rc = ensure_twiddles(ctx, pl.g.D - 1))) return rc;
  size_t B = pick_wave(ctx, pl, n, !device_io, 0);
  if ((rc = ensure_arena(ctx, pl.total))) return rc;
  ctx->have_last = false;
  if (!device_io) {
    // uploads run on the copy stream, one wave ahead of compute
    CU(cudaEventRecord(ctx->ev_free[0], ctx->stream));
    CU(cudaEventRecord(ctx->ev_free[1], ctx->stream));
    if ((rc = stage_upload(ctx, pl, 0, blobs, stride, std::min(B, n)))) return rc;
  }
  int buf = 0;
  for (size_t b0 = 0; b0 < n; b0 += B, buf ^= 1) {
    size_t nb = std::min(B, n - b0);
    Plan w = pl;
    w.B = nb;
    const uint8_t *d_in = blobs + b0 * stride;
    size_t d_stride = stride;
    if (!device_io) {
      if (b0 + B < n &&
          (rc = stage_upload(ctx, pl, buf ^ 1, blobs + (b0 + B) * stride, stride, std::min(B, n - b0 - B))))
        return rc;
      if ((rc = stage_acquire(ctx, pl, buf, &d_in, &d_stride))) return rc;
    }
    if ((rc = lde_wave(ctx, w, d_in, d_stride, device_io ? -1 : buf))) return rc;
    uint8_t *d_roots = at<uint8_t>(ctx, w.o_roots);
    if ((rc = commit_tree(ctx, w, 0, SRC_COLS, nullptr, d_roots, 32))) return rc;
    CU(cudaMemcpyAsync(roots_out + b0 * 32, d_roots, nb * 32, device_io ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                       ctx->stream));
  }
  if (!device_io) CU(cudaStreamSynchronize(ctx->stream));
  return FRIEDA_OK;
}

// ---- FRI commit phase for one wave; leaves all state in the arena ------------------------
int fri_wave(frieda_ctx *ctx, const Plan &w, const uint8_t *d_in, size_t d_stride, const uint64_t *d_seeds,
             int staged_buf = -1, int *error_flag = nullptr) {
  const Geom &g = w.g;
  int rc;
  if ((rc = lde_wave(ctx, w, d_in, d_stride, staged_buf))) return rc;
  Channel *chan = at<Channel>(ctx, w.o_chan);
  uint8_t *roots = at<uint8_t>(ctx, w.o_roots);
  const size_t roots_stride = (size_t)g.n_layers * 32;
  uint32_t layer = 0;
  KL("channel_init", launch_channel_init(ctx->stream, chan, d_seeds, w.B), 1);
  while (layer < g.n_layers && layer_log(g, layer) > TAIL_LOG) {
    int src = layer == 0 ? SRC_COLS : (layer == 1 ? SRC_FOLD_CIRCLE : SRC_FOLD_LINE);
    if ((rc = commit_tree(ctx, w, layer, src, chan, roots + 32 * (size_t)layer, roots_stride))) return rc;
    layer++;
  }
  const uint32_t s = layer;  // first layer handled by the tail (== n_layers: only the last evaluation)
  const uint32_t s_log = s < g.n_layers ? layer_log(g, s) : g.last_log;
  if (s >= 1) {
    const uint32_t src_log = layer_log(g, s - 1);
    KL("fold", launch_fold(ctx->stream, at<uint32_t>(ctx, w.o_cols[s - 1]), w.cols_stride[s - 1], src_log, s - 1 == 0,
                   at<QM31>(ctx, w.o_alpha) + (s - 1), g.n_layers, table(ctx), at<uint32_t>(ctx, w.o_cols[s]),
                   w.cols_stride[s], w.B),
       1);
  }
  TailParams tp;
  std::memset(&tp, 0, sizeof tp);
  for (uint32_t l = 0; l <= g.n_layers; l++) {
    tp.cols[l] = at<uint32_t>(ctx, w.o_cols[l]);
    tp.cols_stride[l] = w.cols_stride[l];
    if (l < g.n_layers) {
      tp.tree[l] = at<uint8_t>(ctx, w.o_tree[l]);
      tp.tree_stride[l] = w.tree_stride[l];
    }
  }
  tp.write_all = w.keep ? 1 : 0;
  tp.start_layer = s;
  tp.start_log = s_log;
  tp.last_log = g.last_log;
  tp.log_last = g.log_last;
  tp.inv_last_n = m31_inv(1u << g.last_log);
  tp.roots = roots;
  tp.roots_stride = roots_stride;
  tp.chan = chan;
  tp.alpha = at<QM31>(ctx, w.o_alpha);
  tp.alpha_stride = g.n_layers;
  tp.last_poly = at<QM31>(ctx, w.o_last);
  tp.error_flag = error_flag ? error_flag : at<int>(ctx, w.o_err);
  tp.tt = table(ctx);
  KL("fri_tail", launch_tail(ctx->stream, tp, w.B), 1);
  return FRIEDA_OK;
}

int fri_commit_impl(frieda_ctx *ctx, const uint8_t *blobs, size_t len, size_t stride, size_t n,
                    const uint64_t *seeds, const frieda_pcs_config *cfg, uint8_t *roots_out,
                    frieda_qm31 *last_poly_out, bool device_io) {
  if (!ctx) return FRIEDA_ERR_ARG;
  if ((!blobs && len) || !cfg || !roots_out || !last_poly_out) return ctx->fail_arg("null pointer");
  if (n == 0) return FRIEDA_OK;
  if (n > 1 && stride < len) return ctx->fail_arg("blob_stride smaller than blob_len");
  CU(cudaSetDevice(ctx->device));
  Plan pl;
  int rc = make_geom_fri(ctx, len, cfg, pl.g);
  if (rc) return rc;
  pl.fri = true;
  pl.keep = ctx->debug_keep;
  pl.levels_cfg = ctx->levels_for(n, pl.g.D);
  if ((rc = ensure_twiddles(ctx, pl.g.D - 1))) return rc;
  size_t B = pick_wave(ctx, pl, n, !device_io, 0);
  if ((rc = ensure_arena(ctx, pl.total))) return rc;
  const Geom &g = pl.g;
  const cudaMemcpyKind out_kind = device_io ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  CU(cudaMemsetAsync(at<int>(ctx, pl.o_err), 0, sizeof(int), ctx->stream));
  if (!device_io) {
    // uploads run on the copy stream, one wave ahead of compute
    CU(cudaEventRecord(ctx->ev_free[0], ctx->stream));
    CU(cudaEventRecord(ctx->ev_free[1], ctx->stream));
    if ((rc = stage_upload(ctx, pl, 0, blobs, stride, std::min(B, n)))) return rc;
  }
  int buf = 0;
  for (size_t b0 = 0; b0 < n; b0 += B, buf ^= 1) {
    size_t nb = std::min(B, n - b0);
    Plan w = pl;
    w.B = nb;
    const uint8_t *d_in = blobs + b0 * stride;
    size_t d_stride = stride;
    const uint64_t *d_seeds = nullptr;
    if (!device_io) {
      if (b0 + B < n &&
          (rc = stage_upload(ctx, pl, buf ^ 1, blobs + (b0 + B) * stride, stride, std::min(B, n - b0 - B))))
        return rc;
      if ((rc = stage_acquire(ctx, pl, buf, &d_in, &d_stride))) return rc;
      if (seeds) {
        CU(cudaMemcpyAsync(at<uint64_t>(ctx, w.o_seeds), seeds + b0, nb * 8, cudaMemcpyHostToDevice, ctx->stream));
        d_seeds = at<uint64_t>(ctx, w.o_seeds);
      }
    } else if (seeds) {
      d_seeds = seeds + b0;
    }
    if ((rc = fri_wave(ctx, w, d_in, d_stride, d_seeds, device_io ? -1 : buf, device_io ? ctx->d_async_err() : nullptr)))
      return rc;
    CU(cudaMemcpyAsync(roots_out + b0 * g.n_layers * 32, at<uint8_t>(ctx, w.o_roots), nb * g.n_layers * 32, out_kind,
                       ctx->stream));
    CU(cudaMemcpyAsync(last_poly_out + (b0 << g.log_last), at<QM31>(ctx, w.o_last), (nb * sizeof(QM31)) << g.log_last,
                       out_kind, ctx->stream));
    ctx->last = w;
    ctx->have_last = true;
  }
  if (!device_io) {
    int err_flag = 0;
    CU(cudaMemcpyAsync(&err_flag, at<int>(ctx, pl.o_err), sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (err_flag) return ctx->fail_arg("reference panics: invalid degree", FRIEDA_ERR_PANIC);
  }
  return FRIEDA_OK;
}

int ensure_gather(frieda_ctx *ctx, int idx, size_t bytes) {
  if (ctx->d_gather_bytes[idx] >= bytes) return FRIEDA_OK;
  if (ctx->d_gather[idx]) {
    CU(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->d_gather[idx]);
    ctx->d_gather[idx] = nullptr;
    ctx->d_gather_bytes[idx] = 0;
  }
  size_t want = bytes + bytes / 4 + 4096;
  CU(cudaMalloc(&ctx->d_gather[idx], want));
  ctx->d_gather_bytes[idx] = want;
  return FRIEDA_OK;
}
int ensure_pinned(frieda_ctx *ctx, int idx, size_t bytes) {
  if (ctx->h_pinned_bytes[idx] >= bytes) return FRIEDA_OK;
  if (ctx->h_pinned[idx]) {
    cudaFreeHost(ctx->h_pinned[idx]);
    ctx->h_pinned[idx] = nullptr;
    ctx->h_pinned_bytes[idx] = 0;
  }
  size_t want = bytes + bytes / 4 + 4096;
  CU(cudaHostAlloc(reinterpret_cast<void **>(&ctx->h_pinned[idx]), want, cudaHostAllocDefault));
  ctx->h_pinned_bytes[idx] = want;
  return FRIEDA_OK;
}

// Host copy of one proof wave's results; two of them alternate so that the Proof objects of wave i are
// assembled on host threads while the GPU already works on wave i + 1.
struct ProveHostBuf {  // every pointer into the wave's pinned readback buffer
  const uint8_t *roots = nullptr;
  const frieda_qm31 *last = nullptr, *evals = nullptr;
  const uint32_t *counts = nullptr, *nuniq = nullptr;
  const unsigned long long *offsets = nullptr, *best = nullptr;
  const int *err_flag = nullptr;
  const uint8_t *fri = nullptr, *hash = nullptr;
};

// Assembles the frieda_proof objects of blobs [b0, b0 + nb) (src/proof.rs:67-76).  Returns false on OOM.
bool assemble_wave(const ProveHostBuf &hb, size_t b0, size_t nb, const Geom &g, uint32_t nq,
                   const frieda_pcs_config *cfg, uint8_t *roots_out, frieda_proof **proofs_out) {
  const uint32_t L = g.n_layers;
  // Phase 1, ONE thread: every allocation of the wave (~47 per proof).  Allocating from several short-lived threads
  // measured 4 ms per 128 proofs on 4 threads and 5.8 ms on 16: each thread grows its own malloc arena (mprotect
  // under the process's mmap lock), and that serialises against the page faults of the copies.
  for (size_t b = 0; b < nb; b++) {
    frieda_proof *pr = (frieda_proof *)std::calloc(1, sizeof(frieda_proof));
    if (!pr) return false;
    proofs_out[b0 + b] = pr;  // (a failed call frees whatever hangs off the proofs it has created: prove_impl)
    pr->pcs_config = *cfg;
    pr->log_size_bound = g.p;
    pr->proof_of_work = hb.best[b];
    pr->n_inner_layers = g.n_inner;
    pr->inner_layers = (frieda_layer_proof *)std::calloc(g.n_inner ? g.n_inner : 1, sizeof(frieda_layer_proof));
    pr->n_last_layer_poly = 1u << g.log_last;
    pr->last_layer_poly = (frieda_qm31 *)std::malloc(sizeof(frieda_qm31) << g.log_last);
    pr->n_evaluations = hb.nuniq[b];
    pr->evaluations = (frieda_qm31 *)std::malloc(sizeof(frieda_qm31) * (hb.nuniq[b] ? hb.nuniq[b] : 1));
    if (!pr->inner_layers || !pr->last_layer_poly || !pr->evaluations) return false;
    for (uint32_t l = 0; l < L; l++) {
      frieda_layer_proof *lp = l == 0 ? &pr->first_layer : &pr->inner_layers[l - 1];
      const size_t ci = (b * L + l) * 2;
      lp->n_fri_witness = hb.counts[ci];
      lp->n_hash_witness = hb.counts[ci + 1];
      lp->n_column_witness = 0;
      lp->fri_witness = (frieda_qm31 *)std::malloc(sizeof(frieda_qm31) * (lp->n_fri_witness ? lp->n_fri_witness : 1));
      lp->hash_witness = (uint8_t *)std::malloc(32 * (size_t)(lp->n_hash_witness ? lp->n_hash_witness : 1));
      lp->column_witness = (uint32_t *)std::malloc(4);
      if (!lp->fri_witness || !lp->hash_witness || !lp->column_witness) return false;
    }
  }
  // Phase 2, a few threads: the copies (130 KB per proof, first touch of the new pages)
  auto fill = [&](size_t b) {
    frieda_proof *pr = proofs_out[b0 + b];
    std::memcpy(pr->last_layer_poly, &hb.last[b << g.log_last], sizeof(frieda_qm31) << g.log_last);
    std::memcpy(pr->evaluations, &hb.evals[b * nq], sizeof(frieda_qm31) * hb.nuniq[b]);
    for (uint32_t l = 0; l < L; l++) {
      frieda_layer_proof *lp = l == 0 ? &pr->first_layer : &pr->inner_layers[l - 1];
      const size_t ci = (b * L + l) * 2;
      std::memcpy(lp->commitment, &hb.roots[(b * L + l) * 32], 32);
      std::memcpy(lp->fri_witness, hb.fri + hb.offsets[ci] * sizeof(QM31), sizeof(QM31) * lp->n_fri_witness);
      std::memcpy(lp->hash_witness, hb.hash + hb.offsets[ci + 1] * 32, 32 * (size_t)lp->n_hash_witness);
    }
    if (roots_out) std::memcpy(roots_out + (b0 + b) * 32, &hb.roots[b * L * 32], 32);
  };
  unsigned nt = (unsigned)std::min<size_t>(std::max(1u, std::thread::hardware_concurrency()),
                                           std::min<size_t>(8, (nb + 31) / 32));
  if (nt <= 1) {
    for (size_t b = 0; b < nb; b++) fill(b);
  } else {
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; t++)
      th.emplace_back([&, t]() {
        for (size_t b = t; b < nb; b += nt) fill(b);
      });
    for (auto &x : th) x.join();
  }
  return true;
}

// ---- prove -------------------------------------------------------------------------------
int prove_impl(frieda_ctx *ctx, const uint8_t *blobs, size_t len, size_t stride, size_t n, const uint64_t *seeds,
               const frieda_pcs_config *cfg, uint8_t *roots_out, frieda_proof **proofs_out) {
  if (!ctx) return FRIEDA_ERR_ARG;
  if ((!blobs && len) || !cfg || !proofs_out) return ctx->fail_arg("null pointer");
  if (n == 0) return FRIEDA_OK;
  if (n > 1 && stride < len) return ctx->fail_arg("blob_stride smaller than blob_len");
  if (cfg->n_queries == 0) return ctx->fail_arg("reference never returns for n_queries == 0", FRIEDA_ERR_PANIC);
  if (cfg->n_queries > 4096) return ctx->fail_arg("n_queries > 4096 is not supported");
  if (cfg->pow_bits > 40) return ctx->fail_arg("pow_bits > 40 is not supported");
  CU(cudaSetDevice(ctx->device));
  for (size_t i = 0; i < n; i++) proofs_out[i] = nullptr;
  StageTrace tr;
  Plan pl;
  int rc = make_geom_fri(ctx, len, cfg, pl.g);
  if (rc) return rc;
  pl.fri = true;
  pl.prove = true;
  pl.keep = true;
  pl.levels_cfg = ctx->levels_for(n, pl.g.D);
  const uint32_t nq = (uint32_t)cfg->n_queries;
  if ((rc = ensure_twiddles(ctx, pl.g.D - 1))) return rc;
  size_t B = pick_wave(ctx, pl, n, true, nq);
  tr.mark("plan");
  if ((rc = ensure_arena(ctx, pl.total))) return rc;
  tr.mark("arena");
  const Geom &g = pl.g;
  const uint32_t L = g.n_layers;
  ProveHostBuf hbuf[2];
  // Waves are pipelined three ways: the NEXT wave's input is uploaded on the copy stream while this one computes; this
  // wave's gathered witnesses (the bulk of a proof: ~130 KB at 64 queries) go back on the copy stream while the next
  // wave computes; and a host thread turns them into Proof objects meanwhile.  The last (or only) wave does its
  // readback on the compute stream and is assembled by the calling thread: a single small proof pays for no thread
  // and no cross-stream hop.
  enum { W_OK = 0, W_OOM = 1, W_DEGREE = 2, W_POW = 3, W_CUDA = 4 };
  struct Workers {
    std::thread th[2];
    bool active[2] = {false, false};
    std::atomic<int> failed{W_OK};
    void join(int i) {
      if (active[i]) {
        th[i].join();
        active[i] = false;
      }
    }
    ~Workers() {
      join(0);
      join(1);
    }
  } workers;
  // checks a wave's flags and assembles its proofs; runs on the calling thread or on a worker
  auto finish_wave = [&workers, &g, nq, cfg, roots_out, proofs_out](const ProveHostBuf &hb, size_t b0, size_t nb) {
    if (*hb.err_flag) {
      workers.failed = W_DEGREE;
      return;
    }
    for (size_t b = 0; b < nb; b++)
      if (hb.best[b] == ~0ull) {
        workers.failed = W_POW;
        return;
      }
    if (!assemble_wave(hb, b0, nb, g, nq, cfg, roots_out, proofs_out)) workers.failed = W_OOM;
  };
  auto fail_code = [&]() -> int {
    workers.join(0);
    workers.join(1);
    const int f = workers.failed;
    if (f == W_OK) return FRIEDA_OK;
    for (size_t i = 0; i < n; i++) {  // a failed call returns no proofs
      if (proofs_out[i]) frieda_proof_free(proofs_out[i]);
      proofs_out[i] = nullptr;
    }
    if (f == W_DEGREE) return ctx->fail_arg("reference panics: invalid degree", FRIEDA_ERR_PANIC);
    if (f == W_POW) return ctx->fail_arg("proof of work search exhausted");
    if (f == W_CUDA) return ctx->fail_arg("proof readback failed on the copy stream", FRIEDA_ERR_CUDA);
    return ctx->fail_arg("out of host memory", FRIEDA_ERR_ALLOC);
  };
  // the waves as a function of their own: whichever way it leaves, the epilogue below joins the workers and, on
  // failure, frees the proofs already assembled (a failed call returns none)
  auto run_waves = [&]() -> int {
  int rc = FRIEDA_OK;
  CU(cudaMemsetAsync(at<int>(ctx, pl.o_err), 0, sizeof(int), ctx->stream));
  // uploads run on the copy stream, one wave ahead of compute
  CU(cudaEventRecord(ctx->ev_free[0], ctx->stream));
  CU(cudaEventRecord(ctx->ev_free[1], ctx->stream));
  if ((rc = stage_upload(ctx, pl, 0, blobs, stride, std::min(B, n)))) return rc;
  size_t wave_index = 0;
  for (size_t b0 = 0; b0 < n; b0 += B, wave_index++) {
    size_t nb = std::min(B, n - b0);
    const int bi = (int)(wave_index & 1);
    const bool last_wave = b0 + B >= n;
    ProveHostBuf &hb = hbuf[bi];
    Plan w = pl;
    w.B = nb;
    const uint8_t *d_in;
    size_t d_stride;
    const uint64_t *d_seeds = nullptr;
    if (!last_wave &&
        (rc = stage_upload(ctx, pl, bi ^ 1, blobs + (b0 + B) * stride, stride, std::min(B, n - b0 - B))))
      return rc;
    if ((rc = stage_acquire(ctx, pl, bi, &d_in, &d_stride))) return rc;
    if (seeds) {
      CU(cudaMemcpyAsync(at<uint64_t>(ctx, w.o_seeds), seeds + b0, nb * 8, cudaMemcpyHostToDevice, ctx->stream));
      d_seeds = at<uint64_t>(ctx, w.o_seeds);
    }
    if ((rc = fri_wave(ctx, w, d_in, d_stride, d_seeds, bi))) return rc;
    tr.mark("fri launches");
    // proof of work (src/proof.rs:58): every blob's minimum nonce, no host round trip
    Channel *chan = at<Channel>(ctx, w.o_chan);
    unsigned long long *best = at<unsigned long long>(ctx, w.o_best);
    CU(cudaMemsetAsync(best, 0xff, nb * 8, ctx->stream));
    const uint64_t limit = (uint64_t)1 << (cfg->pow_bits + 12 > 62 ? 62 : cfg->pow_bits + 12);
    KL("grind", launch_grind(ctx->stream, chan, cfg->pow_bits, limit, best, at<unsigned long long>(ctx, w.o_next), nb), 1);
    // queries + decommitment (src/proof.rs:59-66)
    uint32_t *queries = at<uint32_t>(ctx, w.o_queries);
    uint32_t *nuniq = at<uint32_t>(ctx, w.o_nuniq);
    KL("queries", launch_queries(ctx->stream, chan, best, g.D, nq, queries, nuniq, nb), 1);
    DecommitParams dp;
    std::memset(&dp, 0, sizeof dp);
    dp.queries = queries;
    dp.n_unique = nuniq;
    dp.n_queries = nq;
    dp.n_layers = L;
    dp.D = g.D;
    for (uint32_t l = 0; l < L; l++) {
      dp.cols[l] = at<uint32_t>(ctx, w.o_cols[l]);
      dp.cols_stride[l] = w.cols_stride[l];
      dp.tree[l] = at<uint8_t>(ctx, w.o_tree[l]);
      dp.tree_stride[l] = w.tree_stride[l];
    }
    dp.counts = at<uint32_t>(ctx, w.o_counts);
    dp.offsets = at<unsigned long long>(ctx, w.o_offsets);
    dp.lvl = at<uint32_t>(ctx, w.o_lvl);
    dp.lvl_stride = g.D;
    dp.evals_out = at<QM31>(ctx, w.o_evals);
    unsigned long long *d_totals = at<unsigned long long>(ctx, w.o_totals);
    KL("decommit_count", launch_decommit_count(ctx->stream, dp, nb), 1);
    KL("decommit_scan", launch_decommit_scan(ctx->stream, dp, nb, d_totals), 2);
    unsigned long long totals[2] = {0, 0};
    CU(cudaMemcpyAsync(totals, d_totals, 16, cudaMemcpyDeviceToHost, ctx->stream));
    tr.mark("grind+decommit launches");
    CU(cudaStreamSynchronize(ctx->stream));
    tr.mark("sync 1 (fri+grind+count)");
    // gathered witnesses live in a separate allocation sized by the exact totals; the readback buffer holds the
    // small per-blob results first, then the witnesses
    const size_t fri_bytes = (size_t)totals[0] * sizeof(QM31), hash_bytes = (size_t)totals[1] * 32;
    const size_t fri_pad = align_up(fri_bytes, 256);
    Bump hp;
    const size_t h_roots = hp.take(nb * L * 32), h_last = hp.take((nb * sizeof(frieda_qm31)) << g.log_last);
    const size_t h_evals = hp.take(nb * nq * sizeof(frieda_qm31)), h_counts = hp.take(nb * L * 2 * 4);
    const size_t h_offsets = hp.take(nb * L * 2 * 8), h_nuniq = hp.take(nb * 4), h_best = hp.take(nb * 8);
    const size_t h_err = hp.take(sizeof(int)), h_wit = hp.take(fri_pad + hash_bytes + 256);
    workers.join(bi);  // the wave that used these buffers two iterations ago is read back and assembled
    if (workers.failed) return fail_code();
    if ((rc = ensure_gather(ctx, bi, fri_pad + hash_bytes + 256))) return rc;
    if ((rc = ensure_pinned(ctx, bi, hp.off))) return rc;
    uint8_t *d_gather = ctx->d_gather[bi], *hpin = ctx->h_pinned[bi];
    dp.fri_out = reinterpret_cast<QM31 *>(d_gather);
    dp.hash_out = d_gather + fri_pad;
    ctx->prof_begin("decommit_write");
    cudaError_t le = launch_decommit_write(ctx->stream, dp, nb);
    ctx->prof_end();
    if (le != cudaSuccess) return ctx->fail(le, "launch_decommit_write", __LINE__);
    ctx->launches += 2;
    cudaError_t ce = cudaSuccess;
    auto cp = [&](size_t h_off, const void *src, size_t bytes, cudaStream_t st) {
      if (ce == cudaSuccess && bytes) ce = cudaMemcpyAsync(hpin + h_off, src, bytes, cudaMemcpyDeviceToHost, st);
    };
    cp(h_roots, at<uint8_t>(ctx, w.o_roots), nb * L * 32, ctx->stream);
    cp(h_last, at<QM31>(ctx, w.o_last), (nb * sizeof(frieda_qm31)) << g.log_last, ctx->stream);
    cp(h_evals, dp.evals_out, nb * nq * sizeof(frieda_qm31), ctx->stream);
    cp(h_counts, dp.counts, nb * L * 2 * 4, ctx->stream);
    cp(h_offsets, dp.offsets, nb * L * 2 * 8, ctx->stream);
    cp(h_nuniq, nuniq, nb * 4, ctx->stream);
    cp(h_best, best, nb * 8, ctx->stream);
    cp(h_err, at<int>(ctx, w.o_err), sizeof(int), ctx->stream);
    hb.roots = hpin + h_roots;
    hb.last = reinterpret_cast<const frieda_qm31 *>(hpin + h_last);
    hb.evals = reinterpret_cast<const frieda_qm31 *>(hpin + h_evals);
    hb.counts = reinterpret_cast<const uint32_t *>(hpin + h_counts);
    hb.offsets = reinterpret_cast<const unsigned long long *>(hpin + h_offsets);
    hb.nuniq = reinterpret_cast<const uint32_t *>(hpin + h_nuniq);
    hb.best = reinterpret_cast<const unsigned long long *>(hpin + h_best);
    hb.err_flag = reinterpret_cast<const int *>(hpin + h_err);
    hb.fri = hpin + h_wit;
    hb.hash = hpin + h_wit + fri_pad;
    // (two pieces: the alignment gap between them is never written on the device)
    if (last_wave) {
      cp(h_wit, d_gather, fri_bytes, ctx->stream);
      cp(h_wit + fri_pad, d_gather + fri_pad, hash_bytes, ctx->stream);
      tr.mark("write launch + readback enqueue");
      if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx->stream);
      tr.mark("sync 2 (write+d2h)");
      if (ce != cudaSuccess) return ctx->fail(ce, "proof readback", __LINE__);
      finish_wave(hb, b0, nb);
      tr.mark("assemble");
    } else {
      // witnesses go back on the copy stream, behind the next wave's upload and under its compute
      if (ce == cudaSuccess) ce = cudaEventRecord(ctx->ev_gathered[bi], ctx->stream);
      if (ce == cudaSuccess) ce = cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_gathered[bi], 0);
      cp(h_wit, d_gather, fri_bytes, ctx->copy_stream);
      cp(h_wit + fri_pad, d_gather + fri_pad, hash_bytes, ctx->copy_stream);
      if (ce == cudaSuccess) ce = cudaEventRecord(ctx->ev_readback[bi], ctx->copy_stream);
      if (ce != cudaSuccess) return ctx->fail(ce, "proof readback", __LINE__);
      cudaEvent_t ev = ctx->ev_readback[bi];
      const int device = ctx->device;
      workers.th[bi] = std::thread([&workers, finish_wave, &hb, ev, device, b0, nb]() {
        if (cudaSetDevice(device) != cudaSuccess || cudaEventSynchronize(ev) != cudaSuccess) {
          workers.failed = W_CUDA;
          return;
        }
        finish_wave(hb, b0, nb);
      });
      workers.active[bi] = true;
    }
    ctx->last = w;
    ctx->have_last = true;
  }
  return FRIEDA_OK;
  };
  const int rc_waves = run_waves();
  if (rc_waves != FRIEDA_OK) {
    const std::string keep = ctx->err;
    workers.join(0);
    workers.join(1);
    for (size_t i = 0; i < n; i++) {
      if (proofs_out[i]) frieda_proof_free(proofs_out[i]);
      proofs_out[i] = nullptr;
    }
    ctx->err = keep;
    return rc_waves;
  }
  return fail_code();
}

}  // namespace

// ============================================================================ C ABI
extern "C" {

int frieda_ctx_create(int device, frieda_ctx **out) {
  if (!out) return FRIEDA_ERR_ARG;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    g_create_error = std::string("no usable CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    return FRIEDA_ERR_CUDA;
  }
  if (device < 0 || device >= count) {
    g_create_error = "device index out of range";
    return FRIEDA_ERR_ARG;
  }
  frieda_ctx *ctx = new (std::nothrow) frieda_ctx();
  if (!ctx) return FRIEDA_ERR_ALLOC;
  ctx->device = device;
  if ((e = cudaSetDevice(device)) != cudaSuccess ||
      (e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaEventCreateWithFlags(&ctx->ev_copied[0], cudaEventDisableTiming)) != cudaSuccess ||
      (e = cudaEventCreateWithFlags(&ctx->ev_copied[1], cudaEventDisableTiming)) != cudaSuccess ||
      (e = cudaEventCreateWithFlags(&ctx->ev_free[0], cudaEventDisableTiming)) != cudaSuccess ||
      (e = cudaEventCreateWithFlags(&ctx->ev_free[1], cudaEventDisableTiming)) != cudaSuccess ||
      (e = cudaEventCreateWithFlags(&ctx->ev_gathered[0], cudaEventDisableTiming)) != cudaSuccess ||
      (e = cudaEventCreateWithFlags(&ctx->ev_gathered[1], cudaEventDisableTiming)) != cudaSuccess ||
      (e = cudaEventCreateWithFlags(&ctx->ev_readback[0], cudaEventDisableTiming)) != cudaSuccess ||
      (e = cudaEventCreateWithFlags(&ctx->ev_readback[1], cudaEventDisableTiming)) != cudaSuccess ||
      (e = cudaMalloc(&ctx->d_scratch, 8192)) != cudaSuccess ||
      (e = cudaMemset(ctx->d_scratch, 0, 8192)) != cudaSuccess) {
    g_create_error = std::string("context setup failed: ") + cudaGetErrorString(e);
    cudaGetLastError();
    delete ctx;
    return FRIEDA_ERR_CUDA;
  }
  if (const char *ev = std::getenv("FRIEDA_MERKLE_LEVELS")) {
    int v = std::atoi(ev);
    if (v >= 1 && v <= 10) ctx->merkle_levels_big = (uint32_t)v;
  }
  if (const char *ev = std::getenv("FRIEDA_MERKLE_LATENCY")) ctx->merkle_latency_form = std::atoi(ev) != 0;
  if (const char *ev = std::getenv("FRIEDA_MERKLE_LATENCY_CHUNK")) ctx->merkle_latency_max_chunk = (uint32_t)std::atoi(ev);
  CPoint cur = {host::GEN_X, host::GEN_Y};
  for (int j = 0; j < 31; j++) {
    ctx->gp.g[j] = cur;
    cur = cpoint_add(cur, cur);
  }
  *out = ctx;
  return FRIEDA_OK;
}

void frieda_ctx_destroy(frieda_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  cudaFree(ctx->d_tw);
  cudaFree(ctx->d_itw);
  cudaFree(ctx->d_tw2);
  cudaFree(ctx->arena);
  cudaFree(ctx->d_scratch);
  for (int i = 0; i < 2; i++) {
    cudaFree(ctx->d_gather[i]);
    if (ctx->h_pinned[i]) cudaFreeHost(ctx->h_pinned[i]);
    if (ctx->ev_gathered[i]) cudaEventDestroy(ctx->ev_gathered[i]);
    if (ctx->ev_readback[i]) cudaEventDestroy(ctx->ev_readback[i]);
  }
  for (auto &r : ctx->prof_recs) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  for (auto e : ctx->prof_pool) cudaEventDestroy(e);
  for (int i = 0; i < 2; i++) {
    if (ctx->ev_copied[i]) cudaEventDestroy(ctx->ev_copied[i]);
    if (ctx->ev_free[i]) cudaEventDestroy(ctx->ev_free[i]);
  }
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char *frieda_last_error(const frieda_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int frieda_ctx_device(const frieda_ctx *ctx) { return ctx ? ctx->device : -1; }
int frieda_ctx_fail_arg(frieda_ctx *ctx, const char *msg) { return ctx->fail_arg(msg); }
int frieda_ctx_fail_cuda(frieda_ctx *ctx, int cuda_error, const char *what) {
  return ctx->fail((cudaError_t)cuda_error, what, 0);
}
void frieda_ctx_count_launches(frieda_ctx *ctx, unsigned n) { ctx->launches += n; }
void frieda_ctx_prof_begin(frieda_ctx *ctx, const char *name) { ctx->prof_begin(name); }
void frieda_ctx_prof_end(frieda_ctx *ctx) { ctx->prof_end(); }

int frieda_ctx_set_workspace_limit(frieda_ctx *ctx, size_t bytes) {
  if (!ctx) return FRIEDA_ERR_ARG;
  ctx->ws_limit = bytes;
  return FRIEDA_OK;
}
uint64_t frieda_ctx_launch_count(const frieda_ctx *ctx) { return ctx ? ctx->launches : 0; }
void *frieda_ctx_stream(const frieda_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
int frieda_ctx_set_debug_keep(frieda_ctx *ctx, int on) {
  if (!ctx) return FRIEDA_ERR_ARG;
  ctx->debug_keep = on != 0;
  return FRIEDA_OK;
}

int frieda_ctx_take_error(frieda_ctx *ctx) {
  if (!ctx) return FRIEDA_ERR_ARG;
  CU(cudaSetDevice(ctx->device));
  int flag = 0;
  CU(cudaMemcpyAsync(&flag, ctx->d_async_err(), sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemsetAsync(ctx->d_async_err(), 0, sizeof(int), ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  if (flag) return ctx->fail_arg("reference panics: invalid degree (reported by an asynchronous *_device call)", FRIEDA_ERR_PANIC);
  return FRIEDA_OK;
}

int frieda_ctx_set_profiling(frieda_ctx *ctx, int on) {
  if (!ctx) return FRIEDA_ERR_ARG;
  ctx->profiling = on != 0;
  return FRIEDA_OK;
}
// Synchronises the stream, then writes one line per kernel name: "name launches total_ms\n".
long frieda_ctx_profile_read(frieda_ctx *ctx, char *out, size_t cap, int reset) {
  if (!ctx || !out || cap == 0) return FRIEDA_ERR_ARG;
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->stream));
  struct Acc {
    const char *name;
    unsigned long n;
    double ms;
  };
  std::vector<Acc> acc;
  for (auto &r : ctx->prof_recs) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) {
      cudaGetLastError();
      ms = 0.f;
    }
    size_t i = 0;
    for (; i < acc.size(); i++)
      if (std::strcmp(acc[i].name, r.name) == 0) break;
    if (i == acc.size()) acc.push_back(Acc{r.name, 0, 0.0});
    acc[i].n += 1;
    acc[i].ms += ms;
  }
  std::string text;
  for (auto &a : acc) {
    char line[160];
    std::snprintf(line, sizeof line, "%s %lu %.6f\n", a.name, a.n, a.ms);
    text += line;
  }
  if (reset) {
    for (auto &r : ctx->prof_recs) {
      ctx->prof_pool.push_back(r.a);
      ctx->prof_pool.push_back(r.b);
    }
    ctx->prof_recs.clear();
  }
  if (text.size() + 1 > cap) return ctx->fail_arg("profile buffer too small");
  std::memcpy(out, text.c_str(), text.size() + 1);
  return (long)text.size();
}

int frieda_commit(frieda_ctx *ctx, const uint8_t *data, size_t len, uint32_t log_blowup, uint8_t root_out[32]) {
  return commit_impl(ctx, data, len, len, 1, log_blowup, root_out, false);
}
int frieda_commit_batch(frieda_ctx *ctx, const uint8_t *blobs, size_t blob_len, size_t blob_stride, size_t n,
                        uint32_t log_blowup, uint8_t *roots_out) {
  return commit_impl(ctx, blobs, blob_len, blob_stride, n, log_blowup, roots_out, false);
}
int frieda_commit_batch_device(frieda_ctx *ctx, const uint8_t *d_blobs, size_t blob_len, size_t blob_stride, size_t n,
                               uint32_t log_blowup, uint8_t *d_roots_out) {
  return commit_impl(ctx, d_blobs, blob_len, blob_stride, n, log_blowup, d_roots_out, true);
}

int frieda_fri_n_inner_layers(size_t blob_len, const frieda_pcs_config *cfg) {
  if (!cfg) return FRIEDA_ERR_ARG;
  frieda_ctx tmp;
  Geom g;
  int rc = make_geom_fri(&tmp, blob_len, cfg, g);
  return rc ? rc : (int)g.n_inner;
}
int frieda_fri_commit_batch(frieda_ctx *ctx, const uint8_t *blobs, size_t blob_len, size_t blob_stride, size_t n,
                            const uint64_t *seeds, const frieda_pcs_config *cfg, uint8_t *layer_roots_out,
                            frieda_qm31 *last_poly_out) {
  return fri_commit_impl(ctx, blobs, blob_len, blob_stride, n, seeds, cfg, layer_roots_out, last_poly_out, false);
}
int frieda_fri_commit_batch_device(frieda_ctx *ctx, const uint8_t *d_blobs, size_t blob_len, size_t blob_stride,
                                   size_t n, const uint64_t *d_seeds, const frieda_pcs_config *cfg,
                                   uint8_t *d_layer_roots_out, frieda_qm31 *d_last_poly_out) {
  return fri_commit_impl(ctx, d_blobs, blob_len, blob_stride, n, d_seeds, cfg, d_layer_roots_out, d_last_poly_out,
                         true);
}

int frieda_prove(frieda_ctx *ctx, const uint8_t *data, size_t len, const uint64_t *seed_or_null,
                 const frieda_pcs_config *cfg, uint8_t root_out[32], frieda_proof **proof_out) {
  return frieda_prove_batch(ctx, data, len, len, 1, seed_or_null, cfg, root_out, proof_out);
}
int frieda_prove_batch(frieda_ctx *ctx, const uint8_t *blobs, size_t blob_len, size_t blob_stride, size_t n,
                       const uint64_t *seeds, const frieda_pcs_config *cfg, uint8_t *roots_out,
                       frieda_proof **proofs_out) {
  int rc = prove_impl(ctx, blobs, blob_len, blob_stride, n, seeds, cfg, roots_out, proofs_out);
  if (rc && proofs_out)
    for (size_t i = 0; i < n; i++) {
      frieda_proof_free(proofs_out[i]);
      proofs_out[i] = nullptr;
    }
  return rc;
}

// ---- split blob (config 5) ------------------------------------------------------------------
// Rank `rank` of `world` = 2^g owns the bit-reversed-order index range [rank N/world, (rank+1) N/world)
// of every column = the Merkle subtree (level g, index rank).  The circle FFT runs its largest
// strides first, so that range is computable from the coefficient vector alone (launch_lde with an
// owned range): no exchange of evaluations.  Only the 32-byte subtree roots travel (NCCL
// all-gather in the host layer), then frieda_merkle_combine hashes the top g levels.
// peers != nullptr: the input lies in `world` slices of slice_len bytes behind peer-mapped pointers, and the
// call returns without synchronising the stream.
static int commit_split_local_impl(frieda_ctx *ctx, const uint8_t *data, size_t len, uint32_t log_blowup,
                                   uint32_t rank, uint32_t world, uint8_t *d_subroot_out, bool device_input,
                                   const PeerPtrs *peers = nullptr, size_t slice_len = 0, bool prepare_only = false,
                                   const uint32_t *wait_flags = nullptr, uint32_t part_len = 0, uint32_t epoch = 0,
                                   int *timeout_flag = nullptr) {
  if (!ctx) return FRIEDA_ERR_ARG;
  if ((!data && !peers && len) || !d_subroot_out) return ctx->fail_arg("null pointer");
  if (world == 0 || (world & (world - 1)) || rank >= world) return ctx->fail_arg("world must be a power of two > rank");
  CU(cudaSetDevice(ctx->device));
  Plan pl;
  int rc = make_geom(ctx, len, log_blowup, pl.g);
  if (rc) return rc;
  const Geom &g = pl.g;
  uint32_t gl = 0;
  while ((1u << gl) < world) gl++;
  if (gl > g.D) return ctx->fail_arg("more ranks than evaluation points");
  const uint32_t rlog = g.D - gl;  // log size of the owned range
  if (g.D < 3) return ctx->fail_arg("domain too small to split");
  if (g.p > 15 && rlog < 14) return ctx->fail_arg("owned range too small for this polynomial size");
  if ((rc = ensure_twiddles(ctx, g.D - 1))) return rc;
  // workspace: staged input, coefficients, owned evaluations, truncated tree
  Bump bp;
  size_t o_in = bp.take(align_up(len ? len : 1, 16));
  size_t o_coef = bp.take((size_t)16 << g.p);
  size_t o_eval = bp.take((size_t)16 << rlog);
  const uint32_t lv = ctx->levels_for(1, rlog);
  size_t slots = tree_slots(rlog, false, lv);
  size_t o_tree = bp.take(slots * 32);
  if ((rc = ensure_arena(ctx, bp.off))) return rc;
  ctx->have_last = false;
  if (prepare_only) return FRIEDA_OK;  // everything that can fail on this rank alone has been checked / allocated
  const uint8_t *d_in = data;
  if (!device_input && !peers) {
    uint8_t *stage = at<uint8_t>(ctx, o_in);
    if (len) CU(cudaMemcpyAsync(stage, data, len, cudaMemcpyHostToDevice, ctx->stream));
    d_in = stage;
  }
  uint32_t *coef = at<uint32_t>(ctx, o_coef);
  uint32_t *eval = at<uint32_t>(ctx, o_eval);
  if (peers)
    KL("pack_peers", launch_pack_peers(ctx->stream, *peers, world, rank, slice_len, len, g.n_felts, g.p, coef, wait_flags,
                                       part_len, epoch, timeout_flag),
       1);
  else
    KL("pack", launch_pack(ctx->stream, d_in, len, align_up(len ? len : 1, 16), 1, g.n_felts, g.p, coef), 1);
  LdeRange rg{(size_t)rank << rlog, rlog};
  KL("lde", launch_lde(ctx->stream, coef, eval, g.p, g.beta, 1, g.n_felts, table(ctx), half_initial_point(g), &rg),
     (g.p > 15 ? 2 : 1));
  MerkleBottomParams mp;
  std::memset(&mp, 0, sizeof mp);
  mp.src_cols = eval;
  mp.src_stride = (size_t)4 << rlog;
  mp.tree = at<uint8_t>(ctx, o_tree);
  mp.tree_stride = slots;
  // the top kernel stores the subtree root straight into the caller's slot when that is 16-byte aligned
  const bool direct = (reinterpret_cast<uintptr_t>(d_subroot_out) & 15) == 0;
  if ((rc = run_tree(ctx, SRC_COLS, mp, rlog, lv, false, 1, direct ? d_subroot_out : nullptr, 32, nullptr, nullptr, 0)))
    return rc;
  if (!direct) CU(cudaMemcpyAsync(d_subroot_out, mp.tree + 32, 32, cudaMemcpyDeviceToDevice, ctx->stream));
  if (!peers) CU(cudaStreamSynchronize(ctx->stream));
  return FRIEDA_OK;
}

int frieda_commit_split_local(frieda_ctx *ctx, const uint8_t *data, size_t len, uint32_t log_blowup, uint32_t rank,
                              uint32_t world, uint8_t *d_subroot_out) {
  return commit_split_local_impl(ctx, data, len, log_blowup, rank, world, d_subroot_out, false);
}
int frieda_commit_split_local_device(frieda_ctx *ctx, const uint8_t *d_data, size_t len, uint32_t log_blowup,
                                     uint32_t rank, uint32_t world, uint8_t *d_subroot_out) {
  return commit_split_local_impl(ctx, d_data, len, log_blowup, rank, world, d_subroot_out, true);
}

int frieda_commit_split_local_peers(frieda_ctx *ctx, const uint8_t *const *peer_slices, uint32_t world,
                                    size_t slice_len, size_t len, uint32_t log_blowup, uint32_t rank,
                                    uint8_t *d_subroot_out) {
  if (!ctx) return FRIEDA_ERR_ARG;
  if (!peer_slices) return ctx->fail_arg("null pointer");
  if (world == 0 || world > MAX_PEERS) return ctx->fail_arg("world must be in 1..64");
  if (slice_len == 0 || (slice_len & 15) || (uint64_t)slice_len * world < len)
    return ctx->fail_arg("slice_len must be a multiple of 16 with world * slice_len >= len");
  PeerPtrs pp;
  for (uint32_t r = 0; r < MAX_PEERS; r++) pp.p[r] = r < world ? peer_slices[r] : nullptr;
  for (uint32_t r = 0; r < world; r++)
    if (!pp.p[r]) return ctx->fail_arg("null peer slice");
  return commit_split_local_impl(ctx, nullptr, len, log_blowup, rank, world, d_subroot_out, true, &pp, slice_len);
}

int frieda_commit_split_peers(frieda_ctx *ctx, const uint8_t *data, size_t len, uint32_t log_blowup, uint32_t rank,
                              uint32_t world, uint8_t *const *peer_slices, size_t slice_len,
                              uint8_t *const *peer_roots, uint32_t *const *peer_flags, uint32_t epoch,
                              uint8_t root_out[32]) {
  if (!ctx) return FRIEDA_ERR_ARG;
  // data == nullptr: this rank's slice already lies in peer_slices[rank] (inputs resident in HBM); nothing is uploaded
  if (!peer_slices || !peer_roots || !peer_flags || !root_out) return ctx->fail_arg("null pointer");
  if (world == 0 || (world & (world - 1)) || world > MAX_PEERS || rank >= world)
    return ctx->fail_arg("world must be a power of two <= 64 and > rank");
  if (slice_len == 0 || (slice_len & 15) || (uint64_t)slice_len * world < len)
    return ctx->fail_arg("slice_len must be a multiple of 16 with world * slice_len >= len");
  CU(cudaSetDevice(ctx->device));
  PeerPtrs sl, rt;
  PeerFlags fl;
  for (uint32_t r = 0; r < MAX_PEERS; r++) {
    sl.p[r] = r < world ? peer_slices[r] : nullptr;
    rt.p[r] = r < world ? peer_roots[r] : nullptr;
    fl.p[r] = r < world ? peer_flags[r] : nullptr;
    if (r < world && (!sl.p[r] || !rt.p[r] || !fl.p[r])) return ctx->fail_arg("null peer pointer");
  }
  // Geometry checks, twiddles and workspace BEFORE this rank touches shared state: a rank that fails here has
  // neither overwritten its slice nor entered a barrier its peers would then wait ~20 s on.
  int rc = commit_split_local_impl(ctx, nullptr, len, log_blowup, rank, world, peer_roots[rank], true, &sl, slice_len,
                                   /*prepare_only=*/true);
  if (rc) return rc;
  int *d_timeout = reinterpret_cast<int *>(ctx->d_scratch + 4096);
  CU(cudaMemsetAsync(d_timeout, 0, sizeof(int), ctx->stream));
  // my slice of the input, over my own PCIe link
  const size_t lo = std::min(len, (size_t)rank * slice_len), hi = std::min(len, lo + slice_len);
  Geom g;
  if ((rc = make_geom(ctx, len, log_blowup, g))) return rc;
  const bool pipelined = ((size_t)4 << g.p) >= 4096 && slice_len >= ((size_t)1 << 20) && slice_len < ((size_t)1 << 31);
  if (pipelined) {
    // Upload in PEER_UPLOAD_PARTS parts on the copy stream, each followed by a signal to every rank; the packing
    // kernel on the compute stream waits per part, so it (and the NVLink reads of the early parts) overlap the rest
    // of the upload instead of starting behind a full barrier.  A slice is not overwritten while a peer may still read
    // the previous call's bytes: every rank's packing precedes its arrival at barrier 1 of that call.
    const uint32_t part_len = (uint32_t)align_up((slice_len + PEER_UPLOAD_PARTS - 1) / PEER_UPLOAD_PARTS, 16);
    CU(cudaEventRecord(ctx->ev_free[0], ctx->stream));
    CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_free[0], 0));
    for (uint32_t k = 0; k < PEER_UPLOAD_PARTS; k++) {
      const size_t a = std::min(hi, lo + (size_t)k * part_len), b = std::min(hi, a + part_len);
      if (b > a && data)
        CU(cudaMemcpyAsync(peer_slices[rank] + (a - lo), data + a, b - a, cudaMemcpyHostToDevice, ctx->copy_stream));
      cudaError_t e = launch_peer_signal(ctx->copy_stream, fl, world, rank, 2 + k, epoch);
      if (e != cudaSuccess) return ctx->fail(e, "launch_peer_signal", __LINE__);
      ctx->launches += 1;
    }
    CU(cudaEventRecord(ctx->ev_copied[0], ctx->copy_stream));
    rc = commit_split_local_impl(ctx, nullptr, len, log_blowup, rank, world, peer_roots[rank], true, &sl, slice_len, false,
                                 peer_flags[rank], part_len, epoch, d_timeout);
    CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_copied[0], 0));
  } else {
    if (hi > lo && data) CU(cudaMemcpyAsync(peer_slices[rank], data + lo, hi - lo, cudaMemcpyHostToDevice, ctx->stream));
    KL("peer_barrier", launch_peer_barrier(ctx->stream, fl, world, rank, 0, epoch, d_timeout), 1);
    rc = commit_split_local_impl(ctx, nullptr, len, log_blowup, rank, world, peer_roots[rank], true, &sl, slice_len);
  }
  if (rc) {
    // a launch failed after the peers may already depend on this rank: still arrive at the second barrier so they fail
    // fast on the root they read (this rank reports its own error) instead of spinning until the barrier's timeout
    const std::string keep = ctx->err;
    launch_peer_barrier(ctx->stream, fl, world, rank, 1, epoch, d_timeout);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->copy_stream);
    cudaGetLastError();
    ctx->err = keep;
    return rc;
  }
  KL("peer_barrier", launch_peer_barrier(ctx->stream, fl, world, rank, 1, epoch, d_timeout), 1);
  uint32_t gl = 0;
  while ((1u << gl) < world) gl++;
  uint8_t *tree = ctx->d_scratch;  // heap order: the subtree roots are level gl
  KL("gather_roots", launch_gather_roots(ctx->stream, rt, world, tree), 1);
  KL("merkle_top", launch_merkle_top(ctx->stream, tree, 2 * (size_t)world, gl, 0, nullptr, 0, nullptr, nullptr, 0, 1), 1);
  int timed_out = 0;
  CU(cudaMemcpyAsync(root_out, tree + 32, 32, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(&timed_out, d_timeout, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  if (timed_out) {
    ctx->err = "peer barrier timed out: a rank of the split commit did not arrive";
    return FRIEDA_ERR_CUDA;
  }
  return FRIEDA_OK;
}

int frieda_merkle_combine_peers(frieda_ctx *ctx, const uint8_t *const *peer_roots, uint32_t world,
                                uint8_t root_out[32]) {
  if (!ctx) return FRIEDA_ERR_ARG;
  if (!peer_roots || !root_out) return ctx->fail_arg("null pointer");
  if (world == 0 || (world & (world - 1)) || world > MAX_PEERS) return ctx->fail_arg("world must be a power of two <= 64");
  CU(cudaSetDevice(ctx->device));
  PeerPtrs pp;
  for (uint32_t r = 0; r < MAX_PEERS; r++) pp.p[r] = r < world ? peer_roots[r] : nullptr;
  for (uint32_t r = 0; r < world; r++)
    if (!pp.p[r]) return ctx->fail_arg("null peer root");
  uint32_t gl = 0;
  while ((1u << gl) < world) gl++;
  uint8_t *tree = ctx->d_scratch;  // heap order: the subtree roots are level gl
  KL("gather_roots", launch_gather_roots(ctx->stream, pp, world, tree), 1);
  KL("merkle_top", launch_merkle_top(ctx->stream, tree, 2 * (size_t)world, gl, 0, nullptr, 0, nullptr, nullptr, 0, 1), 1);
  CU(cudaMemcpyAsync(root_out, tree + 32, 32, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return FRIEDA_OK;
}

int frieda_merkle_combine(frieda_ctx *ctx, const uint8_t *d_subroots, uint32_t world, uint8_t root_out[32]) {
  if (!ctx) return FRIEDA_ERR_ARG;
  if (!d_subroots || !root_out) return ctx->fail_arg("null pointer");
  if (world == 0 || (world & (world - 1)) || world > 64) return ctx->fail_arg("world must be a power of two <= 64");
  CU(cudaSetDevice(ctx->device));
  uint32_t gl = 0;
  while ((1u << gl) < world) gl++;
  // scratch tree (heap order): the subtree roots are level gl; d_scratch holds 2 * 64 slots
  uint8_t *tree = ctx->d_scratch;
  CU(cudaMemcpyAsync(tree + (size_t)world * 32, d_subroots, (size_t)world * 32, cudaMemcpyDeviceToDevice, ctx->stream));
  KL("merkle_top", launch_merkle_top(ctx->stream, tree, 2 * (size_t)world, gl, 0, nullptr, 0, nullptr, nullptr, 0, 1), 1);
  CU(cudaMemcpyAsync(root_out, tree + 32, 32, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return FRIEDA_OK;
}


// ---- FRI commit phase of one oversized blob split over GPUs (src/proof.rs:38-57 on a blob that
//      exceeds one GPU) ------------------------------------------------------------------------------------------
// Rank r of world = 2^g owns the bit-reversed index range [r N_l / world, (r+1) N_l / world) of EVERY layer: fold pairs
// (2i, 2i+1) are rank-local, so a layer costs one exchange of `world` 32-byte subtree roots (the caller's all-gather:
// NCCL over NVLink, or peer memory) and every rank then hashes the top g levels and runs the channel step itself.
// Once a rank's share of a layer is down to 2^SPLIT_MIN_LOG points the layer is gathered and finished on every rank
// by the ordinary single-GPU path.  Call order on every rank:
//   begin -> for l in 0 .. n_split-1: { layer(l) -> [all-gather subtree roots] -> combine(l) } -> handoff
//         -> [all-gather local columns] -> finish
namespace {
constexpr uint32_t SPLIT_MIN_LOG = 10;  // smallest rank-local layer that is still committed split

// lays out the unsplit remainder (layers n_split .. n_layers and the last evaluation) behind the split state
void split_layout_remainder(const frieda_ctx::SplitFri &sp, Plan &w, size_t *end_out) {
  const Geom &g = sp.g;
  w = Plan();
  w.g = g;
  w.B = 1;
  w.fri = true;
  w.keep = sp.keep;
  w.levels_cfg = 10;
  Bump fb;
  fb.off = sp.o_full;
  for (uint32_t l = sp.n_split; l <= g.n_layers; l++) {
    const uint32_t lg = l < g.n_layers ? layer_log(g, l) : g.last_log;
    w.cols_stride[l] = (size_t)4 << lg;
    w.o_cols[l] = fb.take((size_t)16 << lg);
  }
  for (uint32_t l = sp.n_split; l < g.n_layers; l++) {
    w.tree_stride[l] = tree_slots(layer_log(g, l), sp.keep, 10);
    w.o_tree[l] = fb.take(w.tree_stride[l] * 32);
  }
  w.o_chan = sp.o_chan;
  w.o_alpha = sp.o_alpha;
  w.o_roots = sp.o_roots;
  w.o_last = sp.o_last;
  w.o_err = sp.o_err;
  if (end_out) *end_out = fb.off + 4096;
}

// peers != nullptr: the blob lies in `world` slices of slice_len bytes behind peer-mapped pointers; this rank uploads
// its own slice (data = the whole blob in host memory, or nullptr when the slices are already resident), a barrier
// kernel (flag channel 7) orders the uploads, and the packing kernel reads every slice in place over NVLink
struct SplitPeerInput {
  PeerPtrs slices;
  PeerFlags flags;
  size_t slice_len;
  uint32_t epoch;
};
int split_begin_impl(frieda_ctx *ctx, const uint8_t *data, size_t len, bool device_input, const uint64_t *seed_or_null,
                     const frieda_pcs_config *cfg, uint32_t rank, uint32_t world, int keep_trees, uint32_t *n_split_out,
                     uint32_t *n_layers_out, uint32_t *handoff_log_out, const SplitPeerInput *peers = nullptr) {
  if (!ctx) return FRIEDA_ERR_ARG;
  ctx->split.active = false;
  if ((!data && len && !peers) || !cfg) return ctx->fail_arg("null pointer");
  if (world == 0 || (world & (world - 1)) || world > MAX_PEERS || rank >= world)
    return ctx->fail_arg("world must be a power of two <= 64 and > rank");
  CU(cudaSetDevice(ctx->device));
  frieda_ctx::SplitFri sp;
  int rc = make_geom_fri(ctx, len, cfg, sp.g);
  if (rc) return rc;
  const Geom &g = sp.g;
  sp.keep = keep_trees != 0;
  sp.cfg = *cfg;
  if (sp.keep && (cfg->n_queries == 0 || cfg->n_queries > 4096 || cfg->pow_bits > 40))
    return ctx->fail_arg("n_queries must be in 1..4096 and pow_bits <= 40 for a proof");
  sp.rank = rank;
  sp.world = world;
  while ((1u << sp.gl) < world) sp.gl++;
  // committed layers whose rank-local share still has 2^SPLIT_MIN_LOG points
  sp.n_split = 0;
  while (sp.n_split < g.n_layers && g.D - sp.n_split >= sp.gl + SPLIT_MIN_LOG) sp.n_split++;
  if (sp.n_split == 0) return ctx->fail_arg("blob too small to split over this many ranks: use frieda_fri_commit_batch");
  const uint32_t rlog = g.D - sp.gl;
  if (g.p > 15 && rlog < 14) return ctx->fail_arg("owned range too small for this polynomial size");
  if ((rc = ensure_twiddles(ctx, g.D - 1))) return rc;
  Bump bp;
  size_t o_in = bp.take(align_up(len ? len : 1, 16));
  sp.o_coef = bp.take((size_t)16 << g.p);
  for (uint32_t l = 0; l < sp.n_split; l++) sp.o_cols[l] = bp.take((size_t)16 << (layer_log(g, l) - sp.gl));
  for (uint32_t l = 0; l < sp.n_split; l++) {
    const uint32_t d = layer_log(g, l) - sp.gl;
    sp.levels_cfg[l] = ctx->levels_for(1, d);
    sp.tree_slots[l] = tree_slots(d, sp.keep, sp.levels_cfg[l]);
    sp.o_tree[l] = bp.take(sp.tree_slots[l] * 32);
  }
  if (sp.keep) {
    sp.o_toptree = bp.take((size_t)sp.n_split * 2 * world * 32);
    sp.o_best = bp.take(8);
    sp.o_next = bp.take(8);
    sp.o_nuniq = bp.take(4);
    sp.o_queries = bp.take((size_t)cfg->n_queries * 4);
  }
  sp.o_chan = bp.take(sizeof(Channel));
  sp.o_alpha = bp.take(g.n_layers * sizeof(QM31));
  sp.o_roots = bp.take((size_t)g.n_layers * 32);
  sp.o_last = bp.take(sizeof(QM31) << g.log_last);
  sp.o_err = bp.take(256);
  sp.o_seed = bp.take(256);
  sp.o_top = bp.take(2 * (size_t)MAX_PEERS * 32);
  sp.o_sub = bp.take(32);
  sp.o_full = bp.off;
  {
    Plan tmp;
    size_t end = 0;
    split_layout_remainder(sp, tmp, &end);
    bp.off = end;
  }
  if ((rc = ensure_arena(ctx, bp.off))) return rc;
  ctx->have_last = false;
  // (the small copies and memsets first: nothing that needs a copy engine may follow the peer barrier below in
  // stream order -- on one GPU the engine queue is shared by the virtual ranks of the tests, and a copy parked behind
  // rank A's barrier would keep rank B's upload from ever starting)
  const uint64_t *d_seed = nullptr;
  if (seed_or_null) {
    CU(cudaMemcpyAsync(at<uint64_t>(ctx, sp.o_seed), seed_or_null, 8, cudaMemcpyHostToDevice, ctx->stream));
    d_seed = at<uint64_t>(ctx, sp.o_seed);
  }
  KL("channel_init", launch_channel_init(ctx->stream, at<Channel>(ctx, sp.o_chan), d_seed, 1), 1);
  CU(cudaMemsetAsync(at<int>(ctx, sp.o_err), 0, sizeof(int), ctx->stream));
  const uint8_t *d_in = data;
  uint32_t *coef = at<uint32_t>(ctx, sp.o_coef);
  if (peers) {
    const size_t lo = std::min(len, (size_t)rank * peers->slice_len), hi = std::min(len, lo + peers->slice_len);
    if (data && hi > lo)
      CU(cudaMemcpyAsync(const_cast<uint8_t *>(peers->slices.p[rank]), data + lo, hi - lo, cudaMemcpyHostToDevice,
                         ctx->stream));
    int *d_timeout = reinterpret_cast<int *>(ctx->d_scratch + 4096);
    CU(cudaMemsetAsync(d_timeout, 0, sizeof(int), ctx->stream));
    sp.peer_timeout_check = true;
    KL("peer_barrier", launch_peer_barrier(ctx->stream, peers->flags, world, rank, 7, peers->epoch, d_timeout), 1);
    KL("pack", launch_pack_peers(ctx->stream, peers->slices, world, rank, peers->slice_len, len, g.n_felts, g.p, coef), 1);
  } else {
    if (!device_input) {
      uint8_t *stage = at<uint8_t>(ctx, o_in);
      if (len) CU(cudaMemcpyAsync(stage, data, len, cudaMemcpyHostToDevice, ctx->stream));
      d_in = stage;
    }
    KL("pack", launch_pack(ctx->stream, d_in, len, align_up(len ? len : 1, 16), 1, g.n_felts, g.p, coef), 1);
  }
  LdeRange rg{(size_t)rank << rlog, rlog};
  KL("lde", launch_lde(ctx->stream, coef, at<uint32_t>(ctx, sp.o_cols[0]), g.p, g.beta, 1, g.n_felts, table(ctx),
                       half_initial_point(g), &rg),
     (g.p > 15 ? 2 : 1));
  sp.next_layer = 0;
  sp.handed_off = false;
  sp.finished = false;
  sp.active = true;
  ctx->split = sp;
  if (n_split_out) *n_split_out = sp.n_split;
  if (n_layers_out) *n_layers_out = g.n_layers;
  if (handoff_log_out) *handoff_log_out = layer_log(g, sp.n_split - 1) - sp.gl - 1;
  return FRIEDA_OK;
}
}  // namespace

int frieda_fri_split_begin(frieda_ctx *ctx, const uint8_t *data, size_t len, const uint64_t *seed_or_null,
                           const frieda_pcs_config *cfg, uint32_t rank, uint32_t world, int keep_trees,
                           uint32_t *n_split_layers_out, uint32_t *n_layers_out, uint32_t *handoff_log_out) {
  return split_begin_impl(ctx, data, len, false, seed_or_null, cfg, rank, world, keep_trees, n_split_layers_out,
                          n_layers_out, handoff_log_out);
}
int frieda_fri_split_begin_device(frieda_ctx *ctx, const uint8_t *d_data, size_t len, const uint64_t *seed_or_null,
                                  const frieda_pcs_config *cfg, uint32_t rank, uint32_t world, int keep_trees,
                                  uint32_t *n_split_layers_out, uint32_t *n_layers_out, uint32_t *handoff_log_out) {
  return split_begin_impl(ctx, d_data, len, true, seed_or_null, cfg, rank, world, keep_trees, n_split_layers_out,
                          n_layers_out, handoff_log_out);
}

// The blob in peer-mapped slices (as frieda_commit_split_peers): this rank uploads slice `rank` of `data` over its own
// PCIe link, and the packing kernel reads all slices where they lie (the input all-gather is its loads over NVLink).
int frieda_fri_split_begin_peers(frieda_ctx *ctx, const uint8_t *data, size_t len, const uint64_t *seed_or_null,
                                 const frieda_pcs_config *cfg, uint32_t rank, uint32_t world, int keep_trees,
                                 uint8_t *const *peer_slices, size_t slice_len, uint32_t *const *peer_flags, uint32_t epoch,
                                 uint32_t *n_split_layers_out, uint32_t *n_layers_out, uint32_t *handoff_log_out) {
  if (!ctx) return FRIEDA_ERR_ARG;
  if (!peer_slices || !peer_flags) return ctx->fail_arg("null pointer");
  if (world == 0 || world > MAX_PEERS || rank >= world) return ctx->fail_arg("world must be a power of two <= 64 and > rank");
  if (slice_len == 0 || (slice_len & 15) || (uint64_t)slice_len * world < len)
    return ctx->fail_arg("slice_len must be a multiple of 16 with world * slice_len >= len");
  SplitPeerInput pi;
  for (uint32_t r = 0; r < MAX_PEERS; r++) {
    pi.slices.p[r] = r < world ? peer_slices[r] : nullptr;
    pi.flags.p[r] = r < world ? peer_flags[r] : nullptr;
    if (r < world && (!pi.slices.p[r] || !pi.flags.p[r])) return ctx->fail_arg("null peer pointer");
  }
  pi.slice_len = slice_len;
  pi.epoch = epoch;
  return split_begin_impl(ctx, data, len, false, seed_or_null, cfg, rank, world, keep_trees, n_split_layers_out,
                          n_layers_out, handoff_log_out, &pi);
}

// Layer `layer` of this rank's range: fold of the previous layer with its alpha (layers >= 1) fused into the leaf
// hashing, rank-local subtree -> 32-byte subtree root (device memory, the all-gather's input).  Asynchronous.
// rank-local part of one split layer: this rank's columns (folded from the previous layer) and its subtree; the
// subtree root goes to d_subroot_out (device; may be peer-visible memory)
static int split_layer_local(frieda_ctx *ctx, uint32_t layer, uint8_t *d_subroot_out) {
  frieda_ctx::SplitFri &sp = ctx->split;
  const Geom &g = sp.g;
  const uint32_t d = layer_log(g, layer) - sp.gl;  // log size of this rank's share
  MerkleBottomParams mp;
  std::memset(&mp, 0, sizeof mp);
  TwiddleTable tt = table(ctx);
  int src = SRC_COLS;
  if (layer == 0) {
    mp.src_cols = at<uint32_t>(ctx, sp.o_cols[0]);
    mp.src_stride = (size_t)4 << d;
  } else {
    src = layer == 1 ? SRC_FOLD_CIRCLE : SRC_FOLD_LINE;
    mp.src_cols = at<uint32_t>(ctx, sp.o_cols[layer - 1]);
    mp.src_stride = (size_t)4 << (d + 1);
    mp.dst_cols = at<uint32_t>(ctx, sp.o_cols[layer]);
    mp.dst_stride = (size_t)4 << d;
    mp.alpha = at<QM31>(ctx, sp.o_alpha) + (layer - 1);
    mp.alpha_stride = g.n_layers;
    // twiddle of pair i of the WHOLE layer: this rank's pairs start at rank * 2^d
    const size_t pair0 = (size_t)sp.rank << d;
    mp.itw_blk = src == SRC_FOLD_CIRCLE ? tt.iblk(1u << (g.D - 2)) + 2 * (pair0 >> 2)
                                        : tt.iblk(1u << layer_log(g, layer)) + pair0;
  }
  mp.tree = at<uint8_t>(ctx, sp.o_tree[layer]);
  mp.tree_stride = sp.tree_slots[layer];
  // the top kernel stores the subtree root straight into d_subroot_out: no copy-engine command sits between a
  // layer's kernels and its barrier (on one GPU, copies of several virtual ranks share one engine queue, and a copy
  // parked behind rank A's barrier would keep rank B from ever reaching it)
  return run_tree(ctx, src, mp, d, sp.levels_cfg[layer], sp.keep, 1, d_subroot_out, 32, nullptr, nullptr, 0);
}
// top log2(world) levels over the `world` subtree roots already placed in slots [world, 2 world) of the layer's top
// tree, then mix_root + draw alpha on this rank's copy of the channel
static int split_layer_top(frieda_ctx *ctx, uint32_t layer, uint8_t *top) {
  frieda_ctx::SplitFri &sp = ctx->split;
  const Geom &g = sp.g;
  KL("merkle_top", launch_merkle_top(ctx->stream, top, 2 * (size_t)sp.world, sp.gl, sp.keep ? 1 : 0,
                                     at<uint8_t>(ctx, sp.o_roots) + 32 * (size_t)layer,
                                     (size_t)g.n_layers * 32, at<Channel>(ctx, sp.o_chan), at<QM31>(ctx, sp.o_alpha) + layer,
                                     g.n_layers, 1),
     1);
  return FRIEDA_OK;
}
static uint8_t *split_top_tree(frieda_ctx *ctx, uint32_t layer) {
  frieda_ctx::SplitFri &sp = ctx->split;
  // heap order: the subtree roots are level gl; kept per layer when a proof follows (the decommitment's top levels)
  return sp.keep ? at<uint8_t>(ctx, sp.o_toptree) + (size_t)layer * 2 * sp.world * 32 : at<uint8_t>(ctx, sp.o_top);
}

int frieda_fri_split_layer(frieda_ctx *ctx, uint32_t layer, uint8_t *d_subroot_out) {
  if (!ctx) return FRIEDA_ERR_ARG;
  frieda_ctx::SplitFri &sp = ctx->split;
  if (!sp.active || sp.handed_off) return ctx->fail_arg("no split FRI commit in progress (call frieda_fri_split_begin)");
  if (!d_subroot_out) return ctx->fail_arg("null pointer");
  if (layer != sp.next_layer || layer >= sp.n_split) return ctx->fail_arg("split layers must be committed in order");
  CU(cudaSetDevice(ctx->device));
  return split_layer_local(ctx, layer, d_subroot_out);
}

// Top log2(world) levels over the gathered subtree roots (d_subroots = world * 32 bytes, rank order, device), then the
// Fiat-Shamir step of the layer (mix_root, draw alpha) on this rank's copy of the channel.  Asynchronous.
int frieda_fri_split_combine(frieda_ctx *ctx, uint32_t layer, const uint8_t *d_subroots) {
  if (!ctx) return FRIEDA_ERR_ARG;
  frieda_ctx::SplitFri &sp = ctx->split;
  if (!sp.active || sp.handed_off) return ctx->fail_arg("no split FRI commit in progress");
  if (!d_subroots) return ctx->fail_arg("null pointer");
  if (layer != sp.next_layer || layer >= sp.n_split) return ctx->fail_arg("split layers must be committed in order");
  CU(cudaSetDevice(ctx->device));
  uint8_t *top = split_top_tree(ctx, layer);
  CU(cudaMemcpyAsync(top + (size_t)sp.world * 32, d_subroots, (size_t)sp.world * 32, cudaMemcpyDeviceToDevice, ctx->stream));
  int rc = split_layer_top(ctx, layer, top);
  if (rc) return rc;
  sp.next_layer = layer + 1;
  return FRIEDA_OK;
}

// All split layers in ONE call per rank, the per-layer exchange done by the library's own kernels over peer-mapped
// memory (NVLink / NVSwitch) instead of a host-driven all-gather per layer: layer l's subtree root is stored into
// slot l of this rank's symmetric roots area (peer_roots[rank], >= 64 slots of 32 bytes), a stream-ordered barrier
// kernel (flag channel 6, epoch * 64 + l) makes it visible, gather_roots_kernel reads the `world` roots where they
// lie, and every rank hashes the top levels and runs the layer's transcript step.  A final barrier keeps a rank from
// overwriting slot l in its next call while a peer still reads this call's.  Asynchronous; a peer that never arrives
// (~20 s) is reported by frieda_fri_split_finish.  epoch: the caller's call counter over these flag arrays, as in
// frieda_commit_split_peers.
int frieda_fri_split_layers_peers(frieda_ctx *ctx, uint8_t *const *peer_roots, uint32_t *const *peer_flags,
                                  uint32_t epoch) {
  if (!ctx) return FRIEDA_ERR_ARG;
  frieda_ctx::SplitFri &sp = ctx->split;
  if (!sp.active || sp.handed_off) return ctx->fail_arg("no split FRI commit in progress (call frieda_fri_split_begin)");
  if (!peer_roots || !peer_flags) return ctx->fail_arg("null pointer");
  if (sp.next_layer != 0) return ctx->fail_arg("split layers were already committed one by one");
  if (sp.world > MAX_PEERS || sp.n_split > 62) return ctx->fail_arg("too many ranks or split layers for the peer form");
  CU(cudaSetDevice(ctx->device));
  PeerFlags fl;
  for (uint32_t r = 0; r < MAX_PEERS; r++) {
    fl.p[r] = r < sp.world ? peer_flags[r] : nullptr;
    if (r < sp.world && (!fl.p[r] || !peer_roots[r])) return ctx->fail_arg("null peer pointer");
  }
  int *d_timeout = reinterpret_cast<int *>(ctx->d_scratch + 4096);
  if (!sp.peer_timeout_check) CU(cudaMemsetAsync(d_timeout, 0, sizeof(int), ctx->stream));  // (begin_peers cleared it)
  sp.peer_timeout_check = true;
  for (uint32_t layer = 0; layer < sp.n_split; layer++) {
    int rc = split_layer_local(ctx, layer, peer_roots[sp.rank] + 32 * (size_t)layer);
    if (rc) return rc;
    KL("peer_barrier", launch_peer_barrier(ctx->stream, fl, sp.world, sp.rank, 6, epoch * 64u + layer, d_timeout), 1);
    PeerPtrs rt;
    for (uint32_t r = 0; r < MAX_PEERS; r++) rt.p[r] = r < sp.world ? peer_roots[r] + 32 * (size_t)layer : nullptr;
    uint8_t *top = split_top_tree(ctx, layer);
    KL("gather_roots", launch_gather_roots(ctx->stream, rt, sp.world, top), 1);
    if ((rc = split_layer_top(ctx, layer, top))) return rc;
    sp.next_layer = layer + 1;
  }
  KL("peer_barrier", launch_peer_barrier(ctx->stream, fl, sp.world, sp.rank, 6, epoch * 64u + 63u, d_timeout), 1);
  return FRIEDA_OK;
}

// After the last split layer: folds it into this rank's share of the first unsplit layer, written to the caller's
// buffer d_cols_local_out (4 columns x 2^handoff_log u32, device; the all-gather's input).  Asynchronous.
int frieda_fri_split_handoff(frieda_ctx *ctx, uint32_t *d_cols_local_out) {
  if (!ctx) return FRIEDA_ERR_ARG;
  frieda_ctx::SplitFri &sp = ctx->split;
  if (!sp.active || sp.handed_off) return ctx->fail_arg("no split FRI commit in progress");
  if (!d_cols_local_out) return ctx->fail_arg("null pointer");
  if (sp.next_layer != sp.n_split) return ctx->fail_arg("split layers are not all committed yet");
  CU(cudaSetDevice(ctx->device));
  const Geom &g = sp.g;
  const uint32_t s = sp.n_split;
  const uint32_t src_log = layer_log(g, s - 1), local_src = src_log - sp.gl;
  KL("fold", launch_fold_range(ctx->stream, at<uint32_t>(ctx, sp.o_cols[s - 1]), src_log, local_src, (size_t)sp.rank << local_src,
                               s - 1 == 0, at<QM31>(ctx, sp.o_alpha) + (s - 1), table(ctx), d_cols_local_out),
     1);
  sp.handed_off = true;
  return FRIEDA_OK;
}

// d_cols_all = world x (4 columns x 2^local_log u32), rank order (the all-gather of the handoff pieces).  Finishes
// FriProver::commit on this rank from the first unsplit layer with the ordinary single-GPU kernels; every rank
// computes the same bytes.  layer_roots_out = n_layers * 32 bytes, last_poly_out = 2^log_last QM31 (host).
int frieda_fri_split_finish(frieda_ctx *ctx, const uint32_t *d_cols_all, uint8_t *layer_roots_out,
                            frieda_qm31 *last_poly_out) {
  if (!ctx) return FRIEDA_ERR_ARG;
  frieda_ctx::SplitFri &sp = ctx->split;
  if (!sp.active || !sp.handed_off) return ctx->fail_arg("frieda_fri_split_handoff has not been called");
  if (!d_cols_all || !layer_roots_out || !last_poly_out) return ctx->fail_arg("null pointer");
  CU(cudaSetDevice(ctx->device));
  const Geom &g = sp.g;
  const uint32_t s = sp.n_split;
  // a one-blob wave whose layers >= s live behind the split state; roots / alpha / channel are the split state's
  split_layout_remainder(sp, sp.rem, nullptr);
  const Plan &w = sp.rem;
  // [rank][column][2^m] -> [column][rank * 2^m ..]
  const uint32_t full_log = s < g.n_layers ? layer_log(g, s) : g.last_log, m = full_log - sp.gl;
  for (uint32_t c = 0; c < 4; c++)
    CU(cudaMemcpy2DAsync(at<uint32_t>(ctx, w.o_cols[s]) + ((size_t)c << full_log), sizeof(uint32_t) << m,
                         d_cols_all + ((size_t)c << m), sizeof(uint32_t) * 4 << m, sizeof(uint32_t) << m, sp.world,
                         cudaMemcpyDeviceToDevice, ctx->stream));
  int rc;
  Channel *chan = at<Channel>(ctx, w.o_chan);
  uint8_t *roots = at<uint8_t>(ctx, w.o_roots);
  const size_t roots_stride = (size_t)g.n_layers * 32;
  uint32_t layer = s;
  while (layer < g.n_layers && layer_log(g, layer) > TAIL_LOG) {
    // the first layer here has its columns already (gathered); the following ones are folded as usual
    int src = layer == s ? SRC_COLS : (layer == 1 ? SRC_FOLD_CIRCLE : SRC_FOLD_LINE);
    if ((rc = commit_tree(ctx, w, layer, src, chan, roots + 32 * (size_t)layer, roots_stride))) return rc;
    layer++;
  }
  const uint32_t t0 = layer;  // first layer handled by the tail
  if (t0 > s) {
    const uint32_t src_log = layer_log(g, t0 - 1);
    KL("fold", launch_fold(ctx->stream, at<uint32_t>(ctx, w.o_cols[t0 - 1]), w.cols_stride[t0 - 1], src_log, t0 - 1 == 0,
                           at<QM31>(ctx, w.o_alpha) + (t0 - 1), g.n_layers, table(ctx), at<uint32_t>(ctx, w.o_cols[t0]),
                           w.cols_stride[t0], 1),
       1);
  }
  TailParams tp;
  std::memset(&tp, 0, sizeof tp);
  for (uint32_t l = s; l <= g.n_layers; l++) {
    tp.cols[l] = at<uint32_t>(ctx, w.o_cols[l]);
    tp.cols_stride[l] = w.cols_stride[l];
    if (l < g.n_layers) {
      tp.tree[l] = at<uint8_t>(ctx, w.o_tree[l]);
      tp.tree_stride[l] = w.tree_stride[l];
    }
  }
  tp.write_all = sp.keep ? 1 : 0;
  tp.start_layer = t0;
  tp.start_log = t0 < g.n_layers ? layer_log(g, t0) : g.last_log;
  tp.last_log = g.last_log;
  tp.log_last = g.log_last;
  tp.inv_last_n = m31_inv(1u << g.last_log);
  tp.roots = roots;
  tp.roots_stride = roots_stride;
  tp.chan = chan;
  tp.alpha = at<QM31>(ctx, w.o_alpha);
  tp.alpha_stride = g.n_layers;
  tp.last_poly = at<QM31>(ctx, w.o_last);
  tp.error_flag = at<int>(ctx, w.o_err);
  tp.tt = table(ctx);
  KL("fri_tail", launch_tail(ctx->stream, tp, 1), 1);
  int err_flag = 0, timed_out = 0;
  if (sp.peer_timeout_check)
    CU(cudaMemcpyAsync(&timed_out, ctx->d_scratch + 4096, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(layer_roots_out, roots, (size_t)g.n_layers * 32, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(last_poly_out, at<QM31>(ctx, w.o_last), sizeof(QM31) << g.log_last, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(&err_flag, at<int>(ctx, w.o_err), sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  sp.finished = true;
  if (!sp.keep) sp.active = false;
  if (timed_out) {
    sp.active = false;
    ctx->err = "peer barrier timed out: a rank of the split FRI commit did not arrive";
    return FRIEDA_ERR_CUDA;
  }
  if (err_flag) {
    sp.active = false;
    return ctx->fail_arg("reference panics: invalid degree", FRIEDA_ERR_PANIC);
  }
  return FRIEDA_OK;
}

// After `finish` of a commit begun with keep_trees: the rest of commit_and_generate_proof (src/proof.rs:58-66) on this
// rank -- proof of work and queries (replicated: same channel, same nonce, same positions on every rank), then THIS
// rank's share of the decommitment: the evaluations, sibling values and tree nodes of split_plan.hpp that it holds,
// gathered by address.  *share_out is malloc'ed (frieda_buffer_free); the shares of all ranks, in rank order, go to
// frieda_fri_split_assemble on whichever rank wants the Proof.  Ends the split state.
int frieda_fri_split_decommit(frieda_ctx *ctx, uint8_t **share_out, size_t *share_len_out) {
  if (!ctx) return FRIEDA_ERR_ARG;
  frieda_ctx::SplitFri &sp = ctx->split;
  if (!share_out || !share_len_out) return ctx->fail_arg("null pointer");
  *share_out = nullptr;
  *share_len_out = 0;
  if (!sp.active || !sp.finished || !sp.keep)
    return ctx->fail_arg("no finished split FRI commit with kept trees (frieda_fri_split_begin with keep_trees)");
  CU(cudaSetDevice(ctx->device));
  sp.active = false;  // whatever happens below, the state is consumed
  const Geom &g = sp.g;
  const uint32_t nq = (uint32_t)sp.cfg.n_queries;
  Channel *chan = at<Channel>(ctx, sp.o_chan);
  unsigned long long *best = at<unsigned long long>(ctx, sp.o_best);
  CU(cudaMemsetAsync(best, 0xff, 8, ctx->stream));
  const uint64_t limit = (uint64_t)1 << (sp.cfg.pow_bits + 12 > 62 ? 62 : sp.cfg.pow_bits + 12);
  KL("grind", launch_grind(ctx->stream, chan, sp.cfg.pow_bits, limit, best, at<unsigned long long>(ctx, sp.o_next), 1), 1);
  uint32_t *d_queries = at<uint32_t>(ctx, sp.o_queries), *d_nuniq = at<uint32_t>(ctx, sp.o_nuniq);
  KL("queries", launch_queries(ctx->stream, chan, best, g.D, nq, d_queries, d_nuniq, 1), 1);
  unsigned long long nonce = 0;
  uint32_t n_unique = 0;
  std::vector<uint32_t> q(nq);
  std::vector<uint8_t> roots((size_t)g.n_layers * 32);
  std::vector<frieda_qm31> last((size_t)1 << g.log_last);
  CU(cudaMemcpyAsync(&nonce, best, 8, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(&n_unique, d_nuniq, 4, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(q.data(), d_queries, (size_t)nq * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(roots.data(), at<uint8_t>(ctx, sp.o_roots), roots.size(), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(last.data(), at<QM31>(ctx, sp.o_last), last.size() * sizeof(frieda_qm31), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  if (nonce == ~0ull) return ctx->fail_arg("proof of work search exhausted");
  if (n_unique == 0 || n_unique > nq) return ctx->fail_arg("query sampling failed");
  // what the proof consists of, and what of it lies on this rank
  std::vector<SplitItem> items;
  std::vector<uint32_t> n_fri, n_hash;
  split_plan(SplitShape{g.D, g.n_layers, sp.n_split, sp.gl}, q.data(), n_unique, items, n_fri, n_hash);
  std::vector<GatherQ> gq;
  std::vector<GatherH> gh;
  for (const SplitItem &it : items) {
    if (it.owner != sp.rank) continue;
    const uint32_t d = g.D - it.layer;
    const bool split = it.layer < sp.n_split;
    if (it.kind == SK_HASH) {
      const uint4 *p;
      if (split && it.level > sp.gl) {  // inside my subtree: local level = level - gl
        const uint32_t ll = it.level - sp.gl;
        p = at<uint4>(ctx, sp.o_tree[it.layer]) + 2 * (((size_t)1 << ll) + (it.index - ((size_t)sp.rank << ll)));
      } else if (split) {               // the replicated top of a split layer's tree
        p = at<uint4>(ctx, sp.o_toptree) + 2 * ((size_t)it.layer * 2 * sp.world + ((size_t)1 << it.level) + it.index);
      } else {                          // unsplit layer: the whole tree is here
        p = at<uint4>(ctx, sp.rem.o_tree[it.layer]) + 2 * (((size_t)1 << it.level) + it.index);
      }
      gh.push_back(GatherH{p});
    } else if (split) {
      const uint32_t ld = d - sp.gl;
      gq.push_back(GatherQ{at<uint32_t>(ctx, sp.o_cols[it.layer]) + (it.index - ((size_t)sp.rank << ld)), 1u << ld});
    } else {
      gq.push_back(GatherQ{at<uint32_t>(ctx, sp.rem.o_cols[it.layer]) + it.index, 1u << d});
    }
  }
  const size_t dq_bytes = align_up(gq.size() * sizeof(GatherQ) + 16, 256), dh_bytes = align_up(gh.size() * sizeof(GatherH) + 16, 256);
  const size_t oq_bytes = align_up(gq.size() * 16 + 16, 256), oh_bytes = gh.size() * 32 + 32;
  int rc = ensure_gather(ctx, 0, dq_bytes + dh_bytes + oq_bytes + oh_bytes);
  if (rc) return rc;
  uint8_t *d = ctx->d_gather[0];
  if (!gq.empty()) CU(cudaMemcpyAsync(d, gq.data(), gq.size() * sizeof(GatherQ), cudaMemcpyHostToDevice, ctx->stream));
  if (!gh.empty()) CU(cudaMemcpyAsync(d + dq_bytes, gh.data(), gh.size() * sizeof(GatherH), cudaMemcpyHostToDevice, ctx->stream));
  KL("gather_items", launch_gather_items(ctx->stream, reinterpret_cast<const GatherQ *>(d), (uint32_t)gq.size(),
                                         reinterpret_cast<const GatherH *>(d + dq_bytes), (uint32_t)gh.size(),
                                         reinterpret_cast<QM31 *>(d + dq_bytes + dh_bytes), d + dq_bytes + dh_bytes + oq_bytes),
     1);
  // the share: header, transcript results, then my items in proof order
  const size_t words = sizeof(SplitShareHeader) / 4 + n_unique + (size_t)g.n_layers * 8 + ((size_t)4 << g.log_last) + 2 +
                       gq.size() * 4 + gh.size() * 8;
  uint32_t *out = (uint32_t *)std::malloc(words * 4);
  if (!out) return ctx->fail_arg("out of host memory", FRIEDA_ERR_ALLOC);
  SplitShareHeader h{SPLIT_SHARE_MAGIC, sp.world, sp.rank, sp.gl, g.D, g.n_layers, sp.n_split, g.p,
                     sp.cfg.log_blowup_factor, sp.cfg.log_last_layer_degree_bound, nq, 0u, sp.cfg.pow_bits,
                     (uint32_t)nonce, (uint32_t)(nonce >> 32), n_unique};
  size_t pos = 0;
  std::memcpy(out, &h, sizeof h);
  pos += sizeof h / 4;
  std::memcpy(out + pos, q.data(), (size_t)n_unique * 4);
  pos += n_unique;
  std::memcpy(out + pos, roots.data(), roots.size());
  pos += roots.size() / 4;
  std::memcpy(out + pos, last.data(), last.size() * sizeof(frieda_qm31));
  pos += last.size() * 4;
  out[pos++] = (uint32_t)gq.size();
  out[pos++] = (uint32_t)gh.size();
  cudaError_t ce = cudaSuccess;
  if (!gq.empty())
    ce = cudaMemcpyAsync(out + pos, d + dq_bytes + dh_bytes, gq.size() * 16, cudaMemcpyDeviceToHost, ctx->stream);
  pos += gq.size() * 4;
  if (ce == cudaSuccess && !gh.empty())
    ce = cudaMemcpyAsync(out + pos, d + dq_bytes + dh_bytes + oq_bytes, gh.size() * 32, cudaMemcpyDeviceToHost, ctx->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx->stream);
  if (ce != cudaSuccess) {
    std::free(out);
    return ctx->fail(ce, "split decommit readback", __LINE__);
  }
  *share_out = reinterpret_cast<uint8_t *>(out);
  *share_len_out = words * 4;
  return FRIEDA_OK;
}

// ---- erasure recovery ----------------------------------------------------------
// block_evals (host): 4 columns x 2^poly_log evaluations = indices [block 2^p, (block+1) 2^p) of each column
// of the committed evaluation (bit-reversed domain order).  Writes the original `len` bytes to data_out.
int frieda_decode_block(frieda_ctx *ctx, const uint32_t *block_evals, size_t len, uint32_t log_blowup, uint32_t block,
                        uint8_t *data_out) {
  if (!ctx) return FRIEDA_ERR_ARG;
  if (!block_evals || (!data_out && len)) return ctx->fail_arg("null pointer");
  CU(cudaSetDevice(ctx->device));
  Geom g;
  int rc = make_geom(ctx, len, log_blowup, g);
  if (rc) return rc;
  if (g.D < 3) return ctx->fail_arg("domain too small to decode");
  if (block >= (1u << log_blowup)) return ctx->fail_arg("block index out of range");
  if ((rc = ensure_twiddles(ctx, g.D - 1))) return rc;
  Bump bp;
  const size_t n_coef = (size_t)4 << g.p;
  size_t o_ev = bp.take(n_coef * 4), o_coef = bp.take(n_coef * 4), o_out = bp.take(len ? len : 1), o_flag = bp.take(256);
  if ((rc = ensure_arena(ctx, bp.off))) return rc;
  ctx->have_last = false;
  CU(cudaMemcpyAsync(at<uint32_t>(ctx, o_ev), block_evals, n_coef * 4, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemsetAsync(at<int>(ctx, o_flag), 0, sizeof(int), ctx->stream));
  KL("decode_block", launch_decode_block(ctx->stream, at<uint32_t>(ctx, o_ev), at<uint32_t>(ctx, o_coef), g.p, g.beta,
                                         block, table(ctx), len, g.n_felts, at<uint8_t>(ctx, o_out),
                                         at<int>(ctx, o_flag)),
     2 + (g.p > 15 ? g.p - 15 : 0));
  int flag = 0;
  if (len) CU(cudaMemcpyAsync(data_out, at<uint8_t>(ctx, o_out), len, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(&flag, at<int>(ctx, o_flag), sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  if (flag) return ctx->fail_arg("the evaluations are not an encoding of `len` bytes");
  return FRIEDA_OK;
}

// Several coset blocks of one evaluation: the data is decoded from the first block in the list that is an encoding of
// `len` bytes (a sampler that collected more than one block can hand over all of them; a corrupted block is skipped).
int frieda_decode_blocks(frieda_ctx *ctx, const uint32_t *block_evals, const uint32_t *block_ids, size_t n_blocks,
                         size_t len, uint32_t log_blowup, uint8_t *data_out, uint32_t *used_out) {
  if (!ctx) return FRIEDA_ERR_ARG;
  if (!block_evals || !block_ids || n_blocks == 0) return ctx->fail_arg("no block to decode from");
  Geom g;
  int rc = make_geom(ctx, len, log_blowup, g);
  if (rc) return rc;
  const size_t words = (size_t)4 << g.p;
  for (size_t k = 0; k < n_blocks; k++) {
    rc = frieda_decode_block(ctx, block_evals + k * words, len, log_blowup, block_ids[k], data_out);
    if (rc == FRIEDA_OK) {
      if (used_out) *used_out = (uint32_t)k;
      return FRIEDA_OK;
    }
    if (rc != FRIEDA_ERR_ARG) return rc;  // CUDA / allocation errors are not "this block is corrupted"
  }
  return ctx->fail_arg("none of the blocks is an encoding of `len` bytes");
}

// ---- standalone passes ------------------------------------------------------------------------
int frieda_pass_pack(frieda_ctx *ctx, const uint8_t *d_blobs, size_t blob_len, size_t blob_stride, size_t n,
                     uint32_t *d_coeffs) {
  if (!ctx || !d_coeffs) return FRIEDA_ERR_ARG;
  CU(cudaSetDevice(ctx->device));
  Geom g;
  g.len = blob_len;
  g.n_felts = (uint32_t)((blob_len * 8 + 29) / 30);
  uint32_t lg = g.n_felts ? host::ceil_log2(g.n_felts) : 0;
  if (lg < 2) lg = 2;
  g.p = lg - 2;
  KL("pack", launch_pack(ctx->stream, d_blobs, blob_len, blob_stride, n, g.n_felts, g.p, d_coeffs), 1);
  return FRIEDA_OK;
}
int frieda_pass_lde(frieda_ctx *ctx, const uint32_t *d_coeffs, uint32_t poly_log, uint32_t log_blowup, size_t n,
                    uint32_t n_felts, uint32_t *d_evals) {
  if (!ctx || !d_coeffs || !d_evals) return FRIEDA_ERR_ARG;
  CU(cudaSetDevice(ctx->device));
  uint32_t D = poly_log + log_blowup;
  if (D == 0 || D > 28) return ctx->fail_arg("bad domain size");
  int rc;
  if (D >= 3 && (rc = ensure_twiddles(ctx, D - 1))) return rc;
  Geom g;
  g.D = D;
  KL("lde", launch_lde(ctx->stream, d_coeffs, d_evals, poly_log, log_blowup, n, n_felts, table(ctx), half_initial_point(g)),
     1);
  return FRIEDA_OK;
}
int frieda_pass_merkle(frieda_ctx *ctx, const uint32_t *d_cols, uint32_t log, size_t n, uint8_t *d_tree,
                       uint8_t *d_roots) {
  if (!ctx || !d_cols || !d_roots) return FRIEDA_ERR_ARG;
  CU(cudaSetDevice(ctx->device));
  if (log > 28) return ctx->fail_arg("bad tree size");
  const bool keep = d_tree != nullptr;
  const uint32_t lv = ctx->levels_for(n, log);
  size_t slots = tree_slots(log, keep, lv);
  uint8_t *tree = d_tree;
  if (!keep) {
    int rc = ensure_arena(ctx, n * slots * 32 + 256);
    if (rc) return rc;
    ctx->have_last = false;
    tree = ctx->arena;
  }
  MerkleBottomParams mp;
  std::memset(&mp, 0, sizeof mp);
  mp.src_cols = d_cols;
  mp.src_stride = (size_t)4 << log;
  mp.tree = tree;
  mp.tree_stride = slots;
  int rc2 = run_tree(ctx, SRC_COLS, mp, log, lv, keep, n, nullptr, 0, nullptr, nullptr, 0);
  if (rc2) return rc2;
  CU(cudaMemcpy2DAsync(d_roots, 32, tree + 32, slots * 32, 32, n, cudaMemcpyDeviceToDevice, ctx->stream));
  return FRIEDA_OK;
}
int frieda_pass_fold(frieda_ctx *ctx, const uint32_t *d_src, uint32_t log, int is_circle, size_t n,
                     const frieda_qm31 *d_alpha, uint32_t *d_dst) {
  if (!ctx || !d_src || !d_alpha || !d_dst) return FRIEDA_ERR_ARG;
  CU(cudaSetDevice(ctx->device));
  if (log < (is_circle ? 3u : 1u) || log > 28) return ctx->fail_arg("bad layer size");
  int rc = ensure_twiddles(ctx, is_circle ? log - 1 : (log < 2 ? 1 : log));
  if (rc) return rc;
  KL("fold", launch_fold(ctx->stream, d_src, (size_t)4 << log, log, is_circle, reinterpret_cast<const QM31 *>(d_alpha), 1,
                 table(ctx), d_dst, (size_t)4 << (log - 1), n),
     1);
  return FRIEDA_OK;
}
int frieda_twiddles(frieda_ctx *ctx, uint32_t k, uint32_t *tw_out, uint32_t *itw_out) {
  if (!ctx || !tw_out || !itw_out) return FRIEDA_ERR_ARG;
  CU(cudaSetDevice(ctx->device));
  if (k > 27) return ctx->fail_arg("twiddle tree too large");
  int rc = ensure_twiddles(ctx, k);
  if (rc) return rc;
  // the tree of half_odds(k) is the tail of the cached (possibly larger) tree
  size_t len = (size_t)1 << k, off = ((size_t)1 << ctx->tw_K) - len;
  CU(cudaMemcpyAsync(tw_out, ctx->d_tw + off, len * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(itw_out, ctx->d_itw + off, len * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return FRIEDA_OK;
}

long frieda_debug_fetch(frieda_ctx *ctx, int what, size_t blob, uint32_t layer, uint32_t level, void *out,
                        size_t cap) {
  if (!ctx || !out) return FRIEDA_ERR_ARG;
  if (!ctx->have_last) return ctx->fail_arg("no resident wave to inspect");
  const Plan &w = ctx->last;
  const Geom &g = w.g;
  if (blob >= w.B) return ctx->fail_arg("blob index outside the last wave");
  const uint8_t *src = nullptr;
  size_t bytes = 0;
  switch (what) {
    case 0:
      src = ctx->arena + w.o_coef + blob * ((size_t)16 << g.p);
      bytes = (size_t)16 << g.p;
      break;
    case 1:
      if (layer > g.n_layers) return ctx->fail_arg("bad layer");
      src = ctx->arena + w.o_cols[layer] + blob * w.cols_stride[layer] * 4;
      bytes = w.cols_stride[layer] * 4;
      break;
    case 2: {
      if (layer >= g.n_layers || !w.keep) return ctx->fail_arg("tree not kept");
      uint32_t d = layer_log(g, layer);
      if (level > d) return ctx->fail_arg("bad level");
      src = ctx->arena + w.o_tree[layer] + (blob * w.tree_stride[layer] + ((size_t)1 << level)) * 32;
      bytes = (size_t)32 << level;
      break;
    }
    case 3:
      if (layer >= g.n_layers) return ctx->fail_arg("bad layer");
      src = ctx->arena + w.o_alpha + (blob * g.n_layers + layer) * sizeof(QM31);
      bytes = sizeof(QM31);
      break;
    case 4:
      src = ctx->arena + w.o_chan + blob * sizeof(Channel);
      bytes = 32;
      break;
    case 5:
      if (!w.prove) return ctx->fail_arg("no proof in the last wave");
      src = ctx->arena + w.o_best + blob * 8;
      bytes = 8;
      break;
    case 6: {
      if (!w.prove) return ctx->fail_arg("no proof in the last wave");
      uint32_t nu = 0;
      CU(cudaMemcpyAsync(&nu, ctx->arena + w.o_nuniq + blob * 4, 4, cudaMemcpyDeviceToHost, ctx->stream));
      CU(cudaStreamSynchronize(ctx->stream));
      src = ctx->arena + w.o_queries + blob * (size_t)w.nq * 4;
      bytes = (size_t)nu * 4;
      break;
    }
    default:
      return ctx->fail_arg("unknown debug selector");
  }
  if (bytes > cap) return ctx->fail_arg("output buffer too small");
  CU(cudaMemcpyAsync(out, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return (long)bytes;
}

}  // extern "C"
