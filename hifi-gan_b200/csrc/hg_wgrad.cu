// hg_wgrad.cu — weight gradient of a (strided / dilated / grouped) Conv1d as a tcgen05 implicit GEMM (sm_100a).
//
// Replaces the `convolution_backward` weight half that torch autograd runs for every conv of the reference's
// training step (UPSTREAM train.py; call sites src/models.py:35-42,63-68,101-104,153-158,208-214 — SURVEY §3.3).
//
//     dW[q][co][ci] = sum_{b, t < t_out}  dy[b, t, co] * xv[b, t + row(q), col(q) + ci]
//
// with xv the same (strided) view of the forward input that hg_conv1d_general_fwd reads and q the packed tap
// index of hg_conv1d_tap_order, so dW comes out in exactly the layout of the forward's packed weight.
//
// GEMM view: the contraction runs over TIME, which is the outer dimension of the channels-last [B][T][C]
// activations, so both operands are "MN-major" for UMMA (instruction-descriptor bits 15/16): a TMA box of
// (64 channels x rows) with SWIZZLE_128B *is* the canonical MN-major SW128 atom stack (8 rows of 128 bytes,
// 8-row groups SBO = 1024 B apart).  A = the forward input x (M = channels), B = dy (N = output channels).
//   * wide layers (cin per group >= 128): an M tile is 128 input channels (two 64-channel slabs, LBO = slab
//     stride) of ONE tap; taps are row-shifted views of the same box (start address + shift * 128 B).
//   * narrow layers (cin per group = 64 / 32): an M tile is 2 / 4 CONSECUTIVE TAPS of the one channel slab —
//     LBO = tap_step rows, i.e. the "next MN atom" of the descriptor is the same slab shifted by one tap.
// Every M tile owns n_tile fp32 columns of TMEM (up to 512 / n_tile tiles accumulate concurrently), a CTA
// walks its share of the (batch, time-chunk) list and finally adds its partial sums into dW with red.global.
// grid = items x splits: item = (N tile / group, channel pair, tap group), split = slice of the time chunks.
#include "hg_common.cuh"

#include <cstdlib>

#include <atomic>

#include "../../include/hifigan_b200.h"

extern std::atomic<int64_t> g_hg_launches;

namespace {

constexpr int kTk = 64;            // time rows per pipeline stage (4 UMMA K steps)
constexpr int kThreads = 192;      // warp 0 TMA, warp 1 MMA, warps 2..5 epilogue
constexpr int kMaxStages = 6;
constexpr int kMaxTapGroups = 48;

struct TapGroup {
  int col;      // channel-coordinate offset of the box in the strided view (rho * C_total)
  int row0;     // view-row offset of the group's first tap relative to the chunk's first output step
  int q0;       // first packed tap
  int ntaps;
};

struct WgradArgs {
  int batch, t_out;
  int chunks_per_b, total_chunks, nsplit;
  int cin_tile, cout_total, n_tile, tiles_n, grouped;
  int wide;            // 1: M tile = 128 channels of one tap; 0: M tile = consecutive taps of one slab
  int atom_w;          // channels per A slab: 64 (SW128) or 32 (SW64)
  int tpm;             // taps per M tile (narrow: 128 / atom_w, wide: 1)
  int mtiles;          // M tiles accumulated concurrently (<= 512 / n_tile)
  int n_cpairs;        // wide: cin_tile / 128, narrow: 1
  int n_tapgroups;
  int tap_step;
  int a_rows;          // rows of the x box
  int b_w;             // channels per dy slab: 64 or 32
  int b_slabs;         // n_tile / b_w
  int stages;
  uint32_t stage_bytes, x_slab_bytes, dy_slab_bytes, dy_off;
  float* dw;
  TapGroup tg[kMaxTapGroups];
};

struct Bars {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t acc_full;
  uint32_t tmem_base;
  uint32_t pad;
};

// MN-major shared-memory descriptor halves: lo = start>>4 | LBO>>4 << 16, hi = SBO>>4 | version | layout
__device__ __forceinline__ uint32_t desc_lo_mn(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr & 0x3FFFFu) >> 4) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}

__device__ __forceinline__ void umma_mn(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                        uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
      "}\n"
      :
      : "r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_dy,
                const WgradArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  Bars* bars = reinterpret_cast<Bars*>(smem + static_cast<uint32_t>(p.stages) * p.stage_bytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // decode the CTA's item: ((split * tiles_n + nt) * n_cpairs + cp) * n_tapgroups + tg
  int item = blockIdx.x;
  const int tgi = item % p.n_tapgroups; item /= p.n_tapgroups;
  const int cp = item % p.n_cpairs; item /= p.n_cpairs;
  const int nt = item % p.tiles_n;
  const int sp = item / p.tiles_n;
  const TapGroup tg = p.tg[tgi];
  const int chan0 = (p.grouped ? nt * p.cin_tile : 0) + cp * 128;
  const int mt_active = p.wide ? tg.ntaps : (tg.ntaps + p.tpm - 1) / p.tpm;
  // this split's slice of the flattened (batch, chunk) list
  const int c_begin = static_cast<int>(static_cast<long long>(p.total_chunks) * sp / p.nsplit);
  const int c_end = static_cast<int>(static_cast<long long>(p.total_chunks) * (sp + 1) / p.nsplit);

  const uint32_t tmem_cols = 512;
  if (warp == 0 && lane == 0) {
    hg::tma_prefetch_desc(&tm_x);
    hg::tma_prefetch_desc(&tm_dy);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < kMaxStages; ++i) {
        hg::mbar_init(&bars->full[i], 1);
        hg::mbar_init(&bars->empty[i], 1);
      }
      hg::mbar_init(&bars->acc_full, 1);
      hg::fence_mbar_init();
    }
    __syncwarp();
    hg::tmem_alloc(&bars->tmem_base, tmem_cols);
  }
  hg::tc_fence_before();
  __syncthreads();
  hg::tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&bars->tmem_base);

  if (c_begin < c_end) {
    if (warp == 0) {
      // ============================ TMA producer ============================
      if (lane == 0) {
        const uint32_t x_slabs = p.wide ? 2u : 1u;
        const uint32_t bytes = x_slabs * static_cast<uint32_t>(p.a_rows) * p.atom_w * 2 +
                               static_cast<uint32_t>(p.b_slabs) * kTk * p.b_w * 2;
        uint32_t slot = 0, phase = 0;
        int b = c_begin / p.chunks_per_b;
        int ck = c_begin - b * p.chunks_per_b;
        for (int c = c_begin; c < c_end; ++c) {
          const int t0 = ck * kTk;
          hg::mbar_wait(&bars->empty[slot], phase ^ 1u);
          hg::mbar_arrive_expect_tx(&bars->full[slot], bytes);
          uint8_t* st = smem + slot * p.stage_bytes;
          for (uint32_t s = 0; s < x_slabs; ++s)
            hg::tma_load_3d(st + s * p.x_slab_bytes, &tm_x, &bars->full[slot],
                            tg.col + chan0 + static_cast<int>(s) * 64, t0 + tg.row0, b);
          for (int s = 0; s < p.b_slabs; ++s)
            hg::tma_load_3d(st + p.dy_off + s * p.dy_slab_bytes, &tm_dy, &bars->full[slot],
                            nt * p.n_tile + s * p.b_w, t0, b);
          if (++slot == static_cast<uint32_t>(p.stages)) { slot = 0; phase ^= 1u; }
          if (++ck == p.chunks_per_b) { ck = 0; ++b; }
        }
      }
      __syncwarp();
    } else if (warp == 1) {
      // ============================ MMA issuer ==============================
      // instruction descriptor: bf16 x bf16 -> fp32, A and B both MN-major (bits 15, 16), M = 128, N = n_tile
      const uint32_t idesc = hg::umma_idesc_bf16(128, static_cast<uint32_t>(p.n_tile)) | (1u << 15) | (1u << 16);
      const uint32_t a_row_bytes = static_cast<uint32_t>(p.atom_w) * 2;
      const uint32_t b_row_bytes = static_cast<uint32_t>(p.b_w) * 2;
      const uint32_t a_hi = hg::umma_desc_hi(8 * a_row_bytes, p.atom_w == 64 ? 2u : 4u);
      const uint32_t b_hi = hg::umma_desc_hi(8 * b_row_bytes, p.b_w == 64 ? 2u : 4u);
      const uint32_t a_lbo = p.wide ? p.x_slab_bytes : static_cast<uint32_t>(p.tap_step) * a_row_bytes;
      const uint32_t a_mt_step = (static_cast<uint32_t>(p.tpm * p.tap_step) * a_row_bytes) >> 4;   // per M tile
      const uint32_t a_kk_step = (16u * a_row_bytes) >> 4;
      const uint32_t b_kk_step = (16u * b_row_bytes) >> 4;
      const bool leader = hg::elect_one();
      uint32_t slot = 0, phase = 0;
      uint32_t accumulate = 0;
      for (int c = c_begin; c < c_end; ++c) {
        hg::mbar_wait(&bars->full[slot], phase);
        hg::tc_fence_after();
        if (leader) {
          const uint32_t st = hg::smem_u32(smem + slot * p.stage_bytes);
          const uint32_t a_lo0 = desc_lo_mn(st, a_lbo);
          const uint32_t b_lo0 = desc_lo_mn(st + p.dy_off, p.dy_slab_bytes);
          for (int kk = 0; kk < kTk / 16; ++kk) {
            uint32_t a_lo = a_lo0 + kk * a_kk_step;
            const uint32_t b_lo = b_lo0 + kk * b_kk_step;
            for (int mt = 0; mt < mt_active; ++mt) {
              umma_mn(tmem_base + mt * p.n_tile, a_lo, a_hi, b_lo, b_hi, idesc, accumulate);
              a_lo += a_mt_step;
            }
            accumulate = 1;
          }
          hg::umma_commit(&bars->empty[slot]);
        }
        accumulate = 1;
        __syncwarp();
        if (++slot == static_cast<uint32_t>(p.stages)) { slot = 0; phase ^= 1u; }
      }
      if (leader) hg::umma_commit(&bars->acc_full);
      __syncwarp();
    } else {
      // ============================ epilogue ================================
      const int quarter = warp & 3;
      const int m = quarter * 32 + lane;            // accumulator row = (atom, channel)
      hg::mbar_wait(&bars->acc_full, 0);
      hg::tc_fence_after();
      const int co0 = nt * p.n_tile;
      for (int mt = 0; mt < mt_active; ++mt) {
        int q, ci;
        if (p.wide) {
          q = tg.q0 + mt;
          ci = cp * 128 + m;
        } else {
          const int a = m / p.atom_w;
          q = tg.q0 + mt * p.tpm + a;
          ci = m - a * p.atom_w;
        }
        const bool valid = q < tg.q0 + tg.ntaps;
        float* dst = p.dw + (static_cast<size_t>(q) * p.cout_total + co0) * p.cin_tile + ci;
        for (int g = 0; g < p.n_tile / 16; ++g) {
          uint32_t raw[16];
          hg::tmem_ld_32x16(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + mt * p.n_tile + g * 16, raw);
          hg::tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int e = 0; e < 16; ++e)
              atomicAdd(dst + static_cast<size_t>(g * 16 + e) * p.cin_tile, __uint_as_float(raw[e]));
          }
        }
      }
    }
  }

  hg::tc_fence_before();
  __syncthreads();
  hg::tc_fence_after();
  if (warp == 1) hg::tmem_dealloc(tmem_base, tmem_cols);
}

int g_sms = 0, g_smem = 0;

int props() {
  if (!g_sms) {
    int dev = 0;
    HG_CHECK_CUDA(cudaGetDevice(&dev));
    HG_CHECK_CUDA(cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev));
    HG_CHECK_CUDA(cudaDeviceGetAttribute(&g_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  }
  static hg::PerDeviceOnce once;
  if (once.need())
    HG_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem));
  return HG_OK;
}

}  // namespace

extern "C" int hg_conv1d_wgrad(const void* x, const void* dy, int batch, int t_in_rows, int c_total, int t_out,
                               int t_out_rows, int groups, int cout, int ktaps, int stride, int dilation,
                               int pad_left, float* dw_packed, int accumulate, void* stream) {
  HG_REQUIRE(x && dy && dw_packed, "hg_conv1d_wgrad: null pointer");
  HG_REQUIRE(batch > 0 && t_out > 0 && t_in_rows > 0 && t_out_rows >= t_out, "hg_conv1d_wgrad: bad sizes");
  HG_REQUIRE(groups >= 1 && c_total % groups == 0 && cout % groups == 0, "hg_conv1d_wgrad: bad groups");
  HG_REQUIRE(stride >= 1 && t_in_rows % stride == 0, "hg_conv1d_wgrad: input rows %d not a multiple of stride %d",
             t_in_rows, stride);
  const int cin_tile = c_total / groups;
  const int n_tile = groups > 1 ? cout / groups
                                : (cout % 256 == 0) ? 256 : (cout % 128 == 0) ? 128 : (cout % 64 == 0) ? 64 : 32;
  HG_REQUIRE(n_tile == 32 || n_tile == 64 || n_tile == 128 || n_tile == 256, "hg_conv1d_wgrad: bad N tile %d", n_tile);
  HG_REQUIRE(cout % n_tile == 0, "hg_conv1d_wgrad: cout %d not a multiple of %d", cout, n_tile);
  HG_REQUIRE(cin_tile == 32 || cin_tile == 64 || cin_tile % 128 == 0,
             "hg_conv1d_wgrad: cin per group must be 32, 64 or a multiple of 128 (got %d)", cin_tile);
  int order[256];
  HG_REQUIRE(ktaps >= 1 && ktaps <= 256, "hg_conv1d_wgrad: bad tap count");
  int rc = hg_conv1d_tap_order(ktaps, stride, pad_left, order);
  if (rc) return rc;
  HG_REQUIRE(stride == 1 || dilation == 1, "hg_conv1d_wgrad: strided convs must have dilation 1");
  rc = props();
  if (rc) return rc;

  WgradArgs p{};
  p.batch = batch; p.t_out = t_out;
  p.chunks_per_b = (t_out + kTk - 1) / kTk;
  p.total_chunks = batch * p.chunks_per_b;
  p.cin_tile = cin_tile; p.cout_total = cout; p.n_tile = n_tile; p.tiles_n = cout / n_tile;
  p.grouped = groups > 1;
  p.wide = cin_tile >= 128;
  p.atom_w = cin_tile == 32 ? 32 : 64;
  p.tpm = p.wide ? 1 : 128 / p.atom_w;
  p.mtiles = 512 / n_tile;
  p.n_cpairs = p.wide ? cin_tile / 128 : 1;
  p.tap_step = stride == 1 ? dilation : 1;
  p.b_w = n_tile == 32 ? 32 : 64;
  p.b_slabs = n_tile / p.b_w;
  // tap groups: walk the residue boxes in packed order (same rule as make_tap_plan of the forward kernel)
  const int taps_per_group = p.mtiles * p.tpm;
  int max_group_span = 1;
  p.n_tapgroups = 0;
  int q = 0;
  while (q < ktaps) {
    // box of packed tap q: residue and first view row
    const int e0 = order[q] * (stride == 1 ? dilation : 1) - pad_left;
    const int rho = stride == 1 ? 0 : ((e0 % stride) + stride) % stride;
    int n = 1;
    while (q + n < ktaps) {
      const int e = order[q + n] - pad_left;
      const int r = stride == 1 ? 0 : ((e % stride) + stride) % stride;
      if (r != rho) break;
      ++n;
    }
    const int row_first = stride == 1 ? e0 : (e0 - rho) / stride;
    for (int o = 0; o < n; o += taps_per_group) {
      HG_REQUIRE(p.n_tapgroups < kMaxTapGroups, "hg_conv1d_wgrad: too many tap groups");
      TapGroup& g = p.tg[p.n_tapgroups++];
      g.col = rho * c_total;
      g.row0 = row_first + o * p.tap_step;
      g.q0 = q + o;
      g.ntaps = (n - o) < taps_per_group ? (n - o) : taps_per_group;
      const int span = p.wide ? g.ntaps : (g.ntaps + p.tpm - 1) / p.tpm * p.tpm;
      if (span > max_group_span) max_group_span = span;
    }
    q += n;
  }
  p.a_rows = kTk + (max_group_span - 1) * p.tap_step;
  HG_REQUIRE(p.a_rows <= 256, "hg_conv1d_wgrad: halo too large for one TMA box (rows=%d)", p.a_rows);
  p.x_slab_bytes = (static_cast<uint32_t>(p.a_rows) * p.atom_w * 2 + 1023u) & ~1023u;
  p.dy_slab_bytes = static_cast<uint32_t>(kTk) * p.b_w * 2;     // 8 KB or 4 KB: multiples of the swizzle period
  p.dy_off = (p.wide ? 2u : 1u) * p.x_slab_bytes;
  p.stage_bytes = (p.dy_off + static_cast<uint32_t>(p.b_slabs) * p.dy_slab_bytes + 1023u) & ~1023u;
  int stages = (g_smem - 1024 - static_cast<int>(sizeof(Bars))) / static_cast<int>(p.stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  HG_REQUIRE(stages >= 2, "hg_conv1d_wgrad: not enough shared memory");
  p.stages = stages;
  const int items = p.tiles_n * p.n_cpairs * p.n_tapgroups;
  // Time splits per item.  A CTA costs (its chunks + an epilogue worth about two chunks: TMEM -> red.global of its
  // tiles) and the grid runs in ceil(CTAs / resident CTAs) rounds, so pick the split that minimises
  // rounds x (chunks / split + 2) instead of a fixed "about two waves" (which left e.g. the 1024 -> 1024 k = 5 layer
  // with 384 CTAs = 2.6 rounds of 148).
  const size_t smem_cta = 1024 + static_cast<size_t>(stages) * p.stage_bytes + sizeof(Bars);
  const int per_sm = smem_cta * 2 + 2048 <= static_cast<size_t>(228 * 1024) ? 2 : 1;
  // Tuning knobs.  HG_WGRAD_FILL = fraction of the resident CTA slots one round may use, HG_WGRAD_EPI = epilogue cost
  // in chunks.  Weight-gradient launches run on low-priority lanes BESIDE the data-gradient chain and other
  // sub-discriminators, so what the step pays for is their SM-time, not their latency: measured on the batch-16 step
  // (ms/step) fill 1.0: 12.88, 0.75: 12.65, 0.5: 12.54, 0.35: 12.46, 0.25: 12.65 (the fixed two waves before: 13.55).
  static const double fill = [] { const char* e = std::getenv("HG_WGRAD_FILL"); return e ? std::atof(e) : 0.4; }();
  static const double epi = [] { const char* e = std::getenv("HG_WGRAD_EPI"); return e ? std::atof(e) : 2.0; }();
  int resident = static_cast<int>(g_sms * per_sm * fill);
  if (resident < 1) resident = 1;
  int nsplit = 1;
  double best = 1e30;
  const int max_split = p.total_chunks < 4 * g_sms ? p.total_chunks : 4 * g_sms;
  for (int cand = 1; cand <= max_split; ++cand) {
    const int ctas = items * cand;
    if (ctas > 4 * resident && cand > 1) break;
    const int rounds = (ctas + resident - 1) / resident;
    const double cost = rounds * (static_cast<double>((p.total_chunks + cand - 1) / cand) + epi);
    if (cost < best - 1e-9) { best = cost; nsplit = cand; }
  }
  p.nsplit = nsplit;
  p.dw = dw_packed;

  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!accumulate)
    HG_CHECK_CUDA(cudaMemsetAsync(dw_packed, 0, static_cast<size_t>(ktaps) * cout * cin_tile * sizeof(float), st));
  CUtensorMap tm_x, tm_dy;
  rc = hg_encode_tmap_bf16_3d(&tm_x, x, static_cast<uint64_t>(stride) * c_total, t_in_rows / stride, batch,
                              static_cast<uint64_t>(stride) * c_total * 2,
                              static_cast<uint64_t>(t_in_rows) * c_total * 2, p.atom_w, p.a_rows, 1, p.atom_w * 2);
  if (rc) return rc;
  // dy: rows >= t_out are outside the tensor map and read as zero (they may hold anything)
  rc = hg_encode_tmap_bf16_3d(&tm_dy, dy, cout, t_out, batch, static_cast<uint64_t>(cout) * 2,
                              static_cast<uint64_t>(t_out_rows) * cout * 2, p.b_w, kTk, 1, p.b_w * 2);
  if (rc) return rc;
  const size_t smem_bytes = 1024 + static_cast<size_t>(stages) * p.stage_bytes + sizeof(Bars);
  wgrad_tc_kernel<<<items * nsplit, kThreads, smem_bytes, st>>>(tm_x, tm_dy, p);
  HG_CHECK_CUDA(cudaGetLastError());
  g_hg_launches.fetch_add(1, std::memory_order_relaxed);
  return HG_OK;
}
