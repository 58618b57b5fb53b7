// hg_conv1d_tc.cu — stride-1 dilated Conv1d as a persistent tcgen05/TMEM implicit GEMM (sm_100a).
//
// Replaces the torch Conv1d / ConvTranspose1d library calls of the reference Generator
// (src/models.py:35-42 ResBlock1, :63-68 ResBlock2, :101 conv_pre, :104 ups via polyphase packing).
//
// GEMM view, per output tile of 128 time steps x n_tile output channels:
//     D[t, co] = sum_{tap j} sum_{chunk c}  A_c[t + j*dil, 0:KC] * W_{j,c}[co, 0:KC]^T
// * activations are channels-last bf16 [B][T][C]; ONE TMA box of (128 + (k-1)*dil) time rows x KC
//   channels is loaded per K chunk and every tap reads it through a row-shifted UMMA descriptor, so
//   the halo is fetched once instead of k times; TMA zero-fills rows outside [0,T) which implements
//   the conv's "same" zero padding.  Row-shifting works because both TMA and UMMA apply the 128B/64B
//   swizzle as a function of the absolute shared-memory address: a start address moved by r rows
//   (r*128 B, any r) with base_offset = 0 and SBO = 8 rows addresses logical rows r..r+127 of the box.
//   Verified on B200 for SW128 and SW64, dilations 1..5, k up to 11 (profiles/r01_bringup.md).
// * weights are bf16 [k][Cout][Cin] (K-major), streamed tap by tap through a multi-stage ring.
// * warp 0 = TMA producer, warp 1 = MMA issuer (one thread) + TMEM owner, warps 2..5 = epilogue.
//   Accumulators are double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
// * epilogue: bias, up to three residual / MRF addends, scale, raw and leaky-relu'd bf16 stores.
#include "hg_common.cuh"

#include <atomic>

extern std::atomic<int64_t> g_hg_launches;

namespace {

constexpr int kTileM = 128;
constexpr int kThreads = 192;
constexpr int kMaxStages = 8;

struct ConvArgs {
  int batch, t, cin, cout;
  int ktaps, dil, pad_left;
  int n_tile, nchunks, a_rows;
  int tiles_t, tiles_n, num_tiles;
  int stages;
  uint32_t a_slot_bytes, w_stage_bytes;
  const float* bias;
  const __nv_bfloat16* res0;
  const __nv_bfloat16* res1;
  const __nv_bfloat16* res2;
  float scale;
  __nv_bfloat16* out_raw;
  __nv_bfloat16* out_act;
  float slope;
};

struct Barriers {
  uint64_t a_full[2];
  uint64_t a_empty[2];
  uint64_t w_full[kMaxStages];
  uint64_t w_empty[kMaxStages];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
};

__device__ __forceinline__ void add_bf16x8(float (&v)[8], const uint4& r) {
  float2 a = hg::unpack_bf16x2(r.x), b = hg::unpack_bf16x2(r.y), c = hg::unpack_bf16x2(r.z),
         d = hg::unpack_bf16x2(r.w);
  v[0] += a.x; v[1] += a.y; v[2] += b.x; v[3] += b.y;
  v[4] += c.x; v[5] += c.y; v[6] += d.x; v[7] += d.y;
}

// KC = channels per K chunk: 64 -> 128-byte rows / SWIZZLE_128B, 32 -> 64-byte rows / SWIZZLE_64B.
template <int KC>
__global__ void __launch_bounds__(kThreads, 1)
conv1d_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                 const ConvArgs p) {
  constexpr uint32_t kRowBytes = KC * 2;
  constexpr uint32_t kLayout = (KC == 64) ? 2u : 4u;  // UMMA layout type: SW128 / SW64
  constexpr uint32_t kSbo = 8 * kRowBytes;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* a_buf = smem;
  uint8_t* w_buf = smem + 2 * p.a_slot_bytes;
  Barriers* bars = reinterpret_cast<Barriers*>(w_buf + p.stages * p.w_stage_bytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t tmem_cols = (2 * p.n_tile <= 32) ? 32u : (2 * p.n_tile <= 64) ? 64u
                             : (2 * p.n_tile <= 128) ? 128u : (2 * p.n_tile <= 256) ? 256u : 512u;

  if (warp == 0 && lane == 0) {
    hg::tma_prefetch_desc(&tm_x);
    hg::tma_prefetch_desc(&tm_w);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < 2; ++i) {
        hg::mbar_init(&bars->a_full[i], 1);
        hg::mbar_init(&bars->a_empty[i], 1);
        hg::mbar_init(&bars->acc_full[i], 1);
        hg::mbar_init(&bars->acc_empty[i], 4);
      }
      for (int i = 0; i < p.stages; ++i) {
        hg::mbar_init(&bars->w_full[i], 1);
        hg::mbar_init(&bars->w_empty[i], 1);
      }
      hg::fence_mbar_init();
    }
    __syncwarp();
    hg::tmem_alloc(&bars->tmem_base, tmem_cols);
  }
  hg::tc_fence_before();
  __syncthreads();
  hg::tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&bars->tmem_base);

  if (warp == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      uint32_t a_it = 0, w_it = 0;
      const uint32_t a_bytes = static_cast<uint32_t>(p.a_rows) * kRowBytes;
      const uint32_t w_bytes = static_cast<uint32_t>(p.n_tile) * kRowBytes;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int nt = tile % p.tiles_n;
        const int rest = tile / p.tiles_n;
        const int tt = rest % p.tiles_t;
        const int b = rest / p.tiles_t;
        const int t0 = tt * kTileM - p.pad_left;
        for (int c = 0; c < p.nchunks; ++c) {
          const uint32_t slot = a_it & 1u;
          hg::mbar_wait(&bars->a_empty[slot], ((a_it >> 1) & 1u) ^ 1u);
          hg::mbar_arrive_expect_tx(&bars->a_full[slot], a_bytes);
          hg::tma_load_3d(a_buf + slot * p.a_slot_bytes, &tm_x, &bars->a_full[slot], c * KC, t0, b);
          ++a_it;
          for (int j = 0; j < p.ktaps; ++j) {
            const uint32_t s = w_it % p.stages;
            const uint32_t ph = (w_it / p.stages) & 1u;
            hg::mbar_wait(&bars->w_empty[s], ph ^ 1u);
            hg::mbar_arrive_expect_tx(&bars->w_full[s], w_bytes);
            hg::tma_load_3d(w_buf + s * p.w_stage_bytes, &tm_w, &bars->w_full[s], c * KC,
                            nt * p.n_tile, j);
            ++w_it;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ============================ MMA issuer ==============================
    if (lane == 0) {
      const uint32_t idesc = hg::umma_idesc_bf16(kTileM, p.n_tile);
      uint32_t a_it = 0, w_it = 0, acc_it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const uint32_t acc = acc_it & 1u;
        hg::mbar_wait(&bars->acc_empty[acc], ((acc_it >> 1) & 1u) ^ 1u);
        hg::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * p.n_tile;
        uint32_t accumulate = 0;
        for (int c = 0; c < p.nchunks; ++c) {
          const uint32_t slot = a_it & 1u;
          hg::mbar_wait(&bars->a_full[slot], (a_it >> 1) & 1u);
          const uint32_t a_base = hg::smem_u32(a_buf + slot * p.a_slot_bytes);
          for (int j = 0; j < p.ktaps; ++j) {
            const uint32_t s = w_it % p.stages;
            hg::mbar_wait(&bars->w_full[s], (w_it / p.stages) & 1u);
            hg::tc_fence_after();
            const uint32_t a_tap = a_base + static_cast<uint32_t>(j * p.dil) * kRowBytes;
            const uint32_t w_base = hg::smem_u32(w_buf + s * p.w_stage_bytes);
#pragma unroll
            for (int kk = 0; kk < KC / 16; ++kk) {
              const uint64_t da = hg::umma_smem_desc(a_tap + kk * 32, kSbo, kLayout, 0);
              const uint64_t db = hg::umma_smem_desc(w_base + kk * 32, kSbo, kLayout, 0);
              hg::umma_bf16_ss(d_tmem, da, db, idesc, accumulate);
              accumulate = 1;
            }
            hg::umma_commit(&bars->w_empty[s]);
            ++w_it;
          }
          hg::umma_commit(&bars->a_empty[slot]);
          ++a_it;
        }
        hg::umma_commit(&bars->acc_full[acc]);
        ++acc_it;
      }
    }
    __syncwarp();
  } else {
    // ============================ epilogue ================================
    const int quarter = warp & 3;  // TMEM lane quarter this warp may read
    const int row = quarter * 32 + lane;
    uint32_t acc_it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int nt = tile % p.tiles_n;
      const int rest = tile / p.tiles_n;
      const int tt = rest % p.tiles_t;
      const int b = rest / p.tiles_t;
      const int t = tt * kTileM + row;
      const uint32_t acc = acc_it & 1u;
      hg::mbar_wait(&bars->acc_full[acc], (acc_it >> 1) & 1u);
      hg::tc_fence_after();
      const bool valid = t < p.t;
      const size_t row_off = (static_cast<size_t>(b) * p.t + (valid ? t : 0)) * p.cout;
      for (int cg = 0; cg < p.n_tile / 32; ++cg) {
        uint32_t raw[32];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                               acc * p.n_tile + cg * 32;
        hg::tmem_ld_32x32(taddr, raw);
        hg::tmem_ld_wait();
        if (valid) {
          const int ch0 = nt * p.n_tile + cg * 32;
          const size_t off = row_off + ch0;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(raw[q * 8 + e]);
            if (p.bias) {
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + ch0 + q * 8));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + ch0 + q * 8 + 4));
              v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
              v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
            }
            if (p.res0) add_bf16x8(v, *reinterpret_cast<const uint4*>(p.res0 + off + q * 8));
            if (p.res1) add_bf16x8(v, *reinterpret_cast<const uint4*>(p.res1 + off + q * 8));
            if (p.res2) add_bf16x8(v, *reinterpret_cast<const uint4*>(p.res2 + off + q * 8));
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] *= p.scale;
            if (p.out_raw) {
              uint4 o;
              o.x = hg::pack_bf16x2(v[0], v[1]); o.y = hg::pack_bf16x2(v[2], v[3]);
              o.z = hg::pack_bf16x2(v[4], v[5]); o.w = hg::pack_bf16x2(v[6], v[7]);
              *reinterpret_cast<uint4*>(p.out_raw + off + q * 8) = o;
            }
            if (p.out_act) {
              uint4 o;
              o.x = hg::pack_bf16x2(hg::lrelu(v[0], p.slope), hg::lrelu(v[1], p.slope));
              o.y = hg::pack_bf16x2(hg::lrelu(v[2], p.slope), hg::lrelu(v[3], p.slope));
              o.z = hg::pack_bf16x2(hg::lrelu(v[4], p.slope), hg::lrelu(v[5], p.slope));
              o.w = hg::pack_bf16x2(hg::lrelu(v[6], p.slope), hg::lrelu(v[7], p.slope));
              *reinterpret_cast<uint4*>(p.out_act + off + q * 8) = o;
            }
          }
        }
      }
      hg::tc_fence_before();
      __syncwarp();
      if (lane == 0) hg::mbar_arrive(&bars->acc_empty[acc]);
      ++acc_it;
    }
  }

  hg::tc_fence_before();
  __syncthreads();
  hg::tc_fence_after();
  if (warp == 1) hg::tmem_dealloc(tmem_base, tmem_cols);
}

int g_num_sms = 0;
int g_max_smem = 0;

int device_props() {
  if (g_num_sms) return HG_OK;
  int dev = 0;
  HG_CHECK_CUDA(cudaGetDevice(&dev));
  HG_CHECK_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  HG_CHECK_CUDA(cudaDeviceGetAttribute(&g_max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  return HG_OK;
}

}  // namespace

extern "C" int hg_conv1d_fwd(const void* x, const void* w_packed, const float* bias, int batch, int t,
                             int cin, int cout, int ktaps, int dilation, int pad_left,
                             const void* res0, const void* res1, const void* res2, float scale,
                             void* out_raw, void* out_act, float act_slope, void* stream) {
  HG_REQUIRE(x && w_packed, "hg_conv1d_fwd: null input");
  HG_REQUIRE(out_raw || out_act, "hg_conv1d_fwd: no output requested");
  HG_REQUIRE(batch > 0 && t > 0, "hg_conv1d_fwd: empty batch/time (%d,%d)", batch, t);
  HG_REQUIRE(cin % 32 == 0 && cin > 0, "hg_conv1d_fwd: cin=%d must be a multiple of 32", cin);
  HG_REQUIRE(cout % 32 == 0 && cout > 0, "hg_conv1d_fwd: cout=%d must be a multiple of 32", cout);
  HG_REQUIRE(ktaps > 0 && dilation > 0, "hg_conv1d_fwd: bad taps/dilation");
  const int a_rows = kTileM + (ktaps - 1) * dilation;
  HG_REQUIRE(a_rows <= 256, "hg_conv1d_fwd: halo too large for one TMA box (rows=%d > 256)", a_rows);
  int rc = device_props();
  if (rc) return rc;

  const int kc = (cin % 64 == 0) ? 64 : 32;
  ConvArgs p{};
  p.batch = batch; p.t = t; p.cin = cin; p.cout = cout;
  p.ktaps = ktaps; p.dil = dilation; p.pad_left = pad_left;
  p.n_tile = (cout % 256 == 0) ? 256 : (cout % 128 == 0) ? 128 : (cout % 64 == 0) ? 64 : 32;
  p.nchunks = cin / kc;
  p.a_rows = a_rows;
  p.tiles_t = (t + kTileM - 1) / kTileM;
  p.tiles_n = cout / p.n_tile;
  p.num_tiles = batch * p.tiles_t * p.tiles_n;
  p.a_slot_bytes = (static_cast<uint32_t>(a_rows) * kc * 2 + 1023u) & ~1023u;
  p.w_stage_bytes = (static_cast<uint32_t>(p.n_tile) * kc * 2 + 1023u) & ~1023u;
  const int budget = g_max_smem - 1024 /*align*/ - static_cast<int>(sizeof(Barriers)) -
                     2 * static_cast<int>(p.a_slot_bytes);
  int stages = budget / static_cast<int>(p.w_stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  HG_REQUIRE(stages >= 2, "hg_conv1d_fwd: not enough shared memory for the weight ring");
  p.stages = stages;
  p.bias = bias;
  p.res0 = static_cast<const __nv_bfloat16*>(res0);
  p.res1 = static_cast<const __nv_bfloat16*>(res1);
  p.res2 = static_cast<const __nv_bfloat16*>(res2);
  p.scale = scale;
  p.out_raw = static_cast<__nv_bfloat16*>(out_raw);
  p.out_act = static_cast<__nv_bfloat16*>(out_act);
  p.slope = act_slope;

  CUtensorMap tm_x, tm_w;
  const int swz = kc * 2;
  rc = hg_encode_tmap_bf16_3d(&tm_x, x, cin, t, batch, static_cast<uint64_t>(cin) * 2,
                              static_cast<uint64_t>(t) * cin * 2, kc, a_rows, 1, swz);
  if (rc) return rc;
  rc = hg_encode_tmap_bf16_3d(&tm_w, w_packed, cin, cout, ktaps, static_cast<uint64_t>(cin) * 2,
                              static_cast<uint64_t>(cout) * cin * 2, kc, p.n_tile, 1, swz);
  if (rc) return rc;

  const size_t smem_bytes = 1024 + 2 * static_cast<size_t>(p.a_slot_bytes) +
                            static_cast<size_t>(stages) * p.w_stage_bytes + sizeof(Barriers);
  const int grid = p.num_tiles < g_num_sms ? p.num_tiles : g_num_sms;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (kc == 64) {
    HG_CHECK_CUDA(cudaFuncSetAttribute(conv1d_tc_kernel<64>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    conv1d_tc_kernel<64><<<grid, kThreads, smem_bytes, st>>>(tm_x, tm_w, p);
  } else {
    HG_CHECK_CUDA(cudaFuncSetAttribute(conv1d_tc_kernel<32>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    conv1d_tc_kernel<32><<<grid, kThreads, smem_bytes, st>>>(tm_x, tm_w, p);
  }
  HG_CHECK_CUDA(cudaGetLastError());
  g_hg_launches.fetch_add(1, std::memory_order_relaxed);
  return HG_OK;
}
