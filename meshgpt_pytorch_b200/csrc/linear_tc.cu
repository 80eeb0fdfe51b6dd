// Dense linears of the encoder on tcgen05 tensor cores with fp32-class accuracy ("split bf16").
//
// Replaces the cuBLAS SGEMMs behind torch.nn.functional.linear in SAGEConv.lin / lin_l / lin_r
// (torch_geometric, called at meshgpt_pytorch.py:764, :769) and project_dim_codebook (:803).
//   C = act([A1 | A2] W^T + bias),  A fp32 (M, K1+K2), W fp32 (N, K)
// Every fp32 operand x is split as x = hi + lo with hi = bf16(x), lo = bf16(x - hi) and the product is
// accumulated in fp32 (TMEM) as  hi.hi + hi.lo + lo.hi  (the lo.lo term, relative 2^-16, is dropped), i.e.
// three bf16 UMMAs per K step: relative error ~1e-5 instead of bf16's 4e-3, at 3/8 of the cost of an
// fp32-exact 3xTF32 scheme.
//
//   * W is split and re-tiled ONCE at load time into the canonical K-major UMMA image (per pass, per
//     32-wide K chunk: hi block then lo block), so the producer thread streams it with one cp.async.bulk
//     per stage;
//   * A is read as fp32 rows, split and written to shared memory in the same canonical layout by eight
//     loader warps (8 rows x 4 sixteen-byte chunks per warp step: coalesced 128 B reads, conflict-free
//     512 B writes);
//   * unit of work = (128-row tile, pass of NT <= 256 output columns); accumulators are double buffered in
//     TMEM so the epilogue (TMEM -> +bias -> ReLU -> global) of one unit overlaps the MMAs of the next.
// Warp roles (512 threads): w0 W producer, w1 MMA issuer, w2 TMEM allocator, w3 idle, w4-7 epilogue,
// w8-15 A loaders.
#include "common.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"

namespace meshtok {

namespace {

using namespace tc;

constexpr int M_TILE = 128;
constexpr int KC = 32;
constexpr int STAGES = 4;
constexpr int NT_MAX = 256;
constexpr int NUM_THREADS = 512;
constexpr int LOADER_WARPS = 8;
constexpr uint32_t SBO = (KC / 8) * 128;                   // 512
constexpr uint32_t A_PART = M_TILE * KC * 2;               // 8 KB (hi or lo)
constexpr uint32_t W_PART_MAX = NT_MAX * KC * 2;           // 16 KB
constexpr uint32_t STAGE_BYTES = 2 * A_PART + 2 * W_PART_MAX;  // 48 KB
constexpr uint32_t BAR_OFF = STAGES * STAGE_BYTES;
constexpr uint32_t SMEM_TOTAL = BAR_OFF + 256;

struct LinArgs {
  const float* A1;
  const float* A2;
  int K1, K2;
  const uint8_t* w_image;
  const float* bias;
  float* C;
  int N, NT, passes;
  const int* m_dev;
  int m_max;
  int relu;
};

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void split_bf16(float x, float y, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
  const float2 hf = __bfloat1622float2(h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(x - hf.x, y - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

__global__ void __launch_bounds__(NUM_THREADS, 1) linear_tc_kernel(LinArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
  uint64_t* full = bars;               // [STAGES]  producer tx + 8 loader warps
  uint64_t* empty = bars + STAGES;     // [STAGES]  MMA commit
  uint64_t* acc_full = empty + STAGES; // [2]
  uint64_t* acc_empty = acc_full + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = warp_id(), lane = lane_id();
  const int M = a.m_dev ? min(*a.m_dev, a.m_max) : a.m_max;
  const int m_tiles = (M + M_TILE - 1) / M_TILE;
  const int units = m_tiles * a.passes;
  const int K = a.K1 + a.K2;
  const int kchunks = K / KC;
  const uint32_t w_part = (uint32_t)a.NT * KC * 2;

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1 + LOADER_WARPS); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== W producer =====
    if (lane == 0) {
      uint32_t it = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int pass = u % a.passes;
        const uint8_t* src = a.w_image + (size_t)pass * kchunks * (2 * w_part);
        for (int kc = 0; kc < kchunks; ++kc, ++it) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&full[s], 2 * w_part);
          bulk_g2s(smem + s * STAGE_BYTES + 2 * A_PART, src + (size_t)kc * (2 * w_part), 2 * w_part, &full[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16_f32(M_TILE, a.NT, /*bf16=*/true);
      const uint32_t s0 = smem_u32(smem);
      uint32_t it = 0, acc_it = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x, ++acc_it) {
        const uint32_t buf = acc_it & 1, aph = (acc_it >> 1) & 1;
        mbar_wait(&acc_empty[buf], aph ^ 1);
        tcgen05_fence_after();
        const uint32_t d = tmem_base + buf * NT_MAX;
        for (int kc = 0; kc < kchunks; ++kc, ++it) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(&full[s], ph);
          tcgen05_fence_after();
          const uint32_t ah = s0 + s * STAGE_BYTES, al = ah + A_PART;
          const uint32_t wh = ah + 2 * A_PART, wl = wh + w_part;
#pragma unroll
          for (int k16 = 0; k16 < KC / 16; ++k16) {
            const uint32_t ko = k16 * 256;
            const uint64_t dah = make_smem_desc(ah + ko, 128, SBO), dal = make_smem_desc(al + ko, 128, SBO);
            const uint64_t dwh = make_smem_desc(wh + ko, 128, SBO), dwl = make_smem_desc(wl + ko, 128, SBO);
            umma_f16(d, dal, dwh, idesc, (kc | k16) != 0 ? 1u : 0u);  // small terms first
            umma_f16(d, dah, dwl, idesc, 1u);
            umma_f16(d, dah, dwh, idesc, 1u);
          }
          umma_commit(&empty[s]);
        }
        umma_commit(&acc_full[buf]);
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ===== epilogue =====
    const int q = warp & 3;
    uint32_t acc_it = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x, ++acc_it) {
      const int m_tile = u / a.passes, pass = u % a.passes;
      const uint32_t buf = acc_it & 1, aph = (acc_it >> 1) & 1;
      mbar_wait(&acc_full[buf], aph);
      tcgen05_fence_after();
      const int row = m_tile * M_TILE + q * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * NT_MAX;
      const int n0 = pass * a.NT;
      float* crow = a.C + (size_t)row * a.N + n0;
      const float* brow = a.bias ? a.bias + n0 : nullptr;
      for (int c = 0; c < a.NT; c += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + c, v);
        tmem_ld_wait();
        if (row < M) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            float4 o;
            float b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
            if (brow) { const float4 bb = __ldg(reinterpret_cast<const float4*>(brow + c + j)); b0 = bb.x; b1 = bb.y; b2 = bb.z; b3 = bb.w; }
            o.x = __uint_as_float(v[j]) + b0; o.y = __uint_as_float(v[j + 1]) + b1;
            o.z = __uint_as_float(v[j + 2]) + b2; o.w = __uint_as_float(v[j + 3]) + b3;
            if (a.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            *reinterpret_cast<float4*>(crow + c + j) = o;
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
  } else if (warp >= 8) {
    // ===== A loaders: fp32 -> (hi, lo) bf16, canonical layout =====
    const int lw = warp - 8;
    uint32_t it = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      const int m_tile = u / a.passes;
      for (int kc = 0; kc < kchunks; ++kc, ++it) {
        const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* sAh = smem + s * STAGE_BYTES;
        uint8_t* sAl = sAh + A_PART;
        const int k0 = kc * KC;
        const float* base;
        int ld, kk;
        if (k0 < a.K1) { base = a.A1; ld = a.K1; kk = k0; }
        else { base = a.A2; ld = a.K2; kk = k0 - a.K1; }
#pragma unroll
        for (int g = lw; g < M_TILE / 8; g += LOADER_WARPS) {
          const int r = g * 8 + (lane & 7);
          const int k8 = lane >> 3;
          const long long grow = (long long)m_tile * M_TILE + r;
          float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
          if (grow < M) {
            const float4* src = reinterpret_cast<const float4*>(base + (size_t)grow * ld + kk + k8 * 8);
            v0 = src[0];
            v1 = src[1];
          }
          uint4 hi, lo;
          split_bf16(v0.x, v0.y, hi.x, lo.x);
          split_bf16(v0.z, v0.w, hi.y, lo.y);
          split_bf16(v1.x, v1.y, hi.z, lo.z);
          split_bf16(v1.z, v1.w, hi.w, lo.w);
          const uint32_t off = canon_offset(r, k8 * 8, SBO);
          *reinterpret_cast<uint4*>(sAh + off) = hi;
          *reinterpret_cast<uint4*>(sAl + off) = lo;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[s]);
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// W (N, K) fp32 -> image [pass][kchunk]{hi (NT x 32), lo (NT x 32)} canonical K-major, SBO 512.
__global__ void __launch_bounds__(256) build_w_image_kernel(const float* __restrict__ W, int N, int K, int NT, uint8_t* __restrict__ image) {
  const long long total = (long long)N * (K / 8);
  const int kchunks = K / KC;
  const uint32_t w_part = (uint32_t)NT * KC * 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / (K / 8));
    const int k8 = (int)(i % (K / 8));
    const float4* src = reinterpret_cast<const float4*>(W + (size_t)n * K + k8 * 8);
    const float4 v0 = src[0], v1 = src[1];
    uint4 hi, lo;
    split_bf16(v0.x, v0.y, hi.x, lo.x);
    split_bf16(v0.z, v0.w, hi.y, lo.y);
    split_bf16(v1.x, v1.y, hi.z, lo.z);
    split_bf16(v1.z, v1.w, hi.w, lo.w);
    const int pass = n / NT, r = n % NT;
    const int kc = (k8 * 8) / KC, kin = (k8 * 8) % KC;
    const size_t base = ((size_t)pass * kchunks + kc) * (2 * (size_t)w_part);
    const uint32_t off = canon_offset(r, kin, SBO);
    *reinterpret_cast<uint4*>(image + base + off) = hi;
    *reinterpret_cast<uint4*>(image + base + w_part + off) = lo;
  }
}

}  // namespace

int linear_tc_plan(int N, int K1, int K2, int* passes, int* NT) {
  const int K = K1 + K2;
  if (N <= 0 || K1 <= 0 || K1 % KC != 0 || K2 % KC != 0) {
    set_error("tcgen05 linear: input widths must be multiples of %d (got %d, %d); use the SIMT path", KC, K1, K2);
    return MESHTOK_ERR_UNSUPPORTED;
  }
  const int p = ceil_div(N, NT_MAX);
  if (N % p != 0 || (N / p) % 16 != 0) {
    set_error("tcgen05 linear: output width %d does not split into %d passes of a multiple of 16 columns; use the SIMT path", N, p);
    return MESHTOK_ERR_UNSUPPORTED;
  }
  (void)K;
  *passes = p;
  *NT = N / p;
  return MESHTOK_OK;
}

size_t linear_tc_image_bytes(int N, int K) { return (size_t)N * K * 4; }

int launch_linear_tc_build_image(const float* W, int N, int K, void* image, cudaStream_t stream) {
  int passes, NT;
  int rc = linear_tc_plan(N, K, 0, &passes, &NT);
  if (rc != MESHTOK_OK) return rc;
  const long long total = (long long)N * (K / 8);
  const long long want = (total + 255) / 256;
  const unsigned blocks = (unsigned)(want < 148 * 8 ? want : 148 * 8);
  build_w_image_kernel<<<blocks, 256, 0, stream>>>(W, N, K, NT, static_cast<uint8_t*>(image));
  MT_CHECK_LAUNCH();
  return MESHTOK_OK;
}

int launch_linear_tc(const float* A1, int K1, const float* A2, int K2, const void* w_image, const float* bias, float* C, int N,
                     const int* m_dev, int m_max, bool relu, cudaStream_t stream) {
  if (m_max == 0) return MESHTOK_OK;
  MT_CHECK_ARG(w_image != nullptr, "tcgen05 linear: weight image is null");
  LinArgs a;
  a.A1 = A1; a.A2 = A2; a.K1 = K1; a.K2 = K2; a.w_image = static_cast<const uint8_t*>(w_image); a.bias = bias; a.C = C;
  a.N = N; a.m_dev = m_dev; a.m_max = m_max; a.relu = relu ? 1 : 0;
  int rc = linear_tc_plan(N, K1, K2, &a.passes, &a.NT);
  if (rc != MESHTOK_OK) return rc;
  static bool configured = false;
  if (!configured) {
    MT_CHECK_CUDA(cudaFuncSetAttribute(linear_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_TOTAL));
    configured = true;
  }
  const int units = ceil_div(m_max, M_TILE) * a.passes;
  const int grid = min(units, 148);
  linear_tc_kernel<<<grid, NUM_THREADS, SMEM_TOTAL, stream>>>(a);
  MT_CHECK_LAUNCH();
  return MESHTOK_OK;
}

}  // namespace meshtok
