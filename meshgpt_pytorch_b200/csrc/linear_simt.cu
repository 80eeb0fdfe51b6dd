// fp32 SIMT linear: C = act([A1 | A2] W^T + bias).  The exact-precision path used by the parity tests
// and for shapes the tcgen05 kernel does not cover.  Replaces torch.nn.functional.linear
// (cuBLAS SGEMM) inside SAGEConv and at meshgpt_pytorch.py:741, :803.
// 64x64 output tile, 16-deep K slices, 256 threads x (4x4) register micro-tiles.
#include "common.cuh"
#include "kernels.cuh"

namespace meshtok {

namespace {

constexpr int BM = 64, BN = 64, BK = 16;

template <bool RELU>
__global__ void __launch_bounds__(256) linear_simt_kernel(const float* __restrict__ A1, int K1, const float* __restrict__ A2,
                                                          int K2, const float* __restrict__ W, const float* __restrict__ bias,
                                                          float* __restrict__ C, int N, const int* __restrict__ m_dev, int m_max) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Ws[BK][BN + 4];
  const int M = m_dev ? min(*m_dev, m_max) : m_max;
  const int m0 = blockIdx.x * BM;
  if (m0 >= M) return;
  const int n0 = blockIdx.y * BN;
  const int K = K1 + K2;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int lr = tid >> 2;          // 0..63: tile row (A) / tile col (W) this thread loads
  const int lk = (tid & 3) * 4;     // k offset inside the slice

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
    // A slice (K1 and K2 are multiples of BK, so a slice never straddles the two sources)
    float4 av = make_float4(0.f, 0.f, 0.f, 0.f);
    const int am = m0 + lr;
    if (am < M) {
      const float* srcp = (k0 < K1) ? (A1 + (size_t)am * K1 + k0 + lk) : (A2 + (size_t)am * K2 + (k0 - K1) + lk);
      av = *reinterpret_cast<const float4*>(srcp);
    }
    float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
    const int wn = n0 + lr;
    if (wn < N) wv = __ldg(reinterpret_cast<const float4*>(W + (size_t)wn * K + k0 + lk));
    As[lk + 0][lr] = av.x; As[lk + 1][lr] = av.y; As[lk + 2][lr] = av.z; As[lk + 3][lr] = av.w;
    Ws[lk + 0][lr] = wv.x; Ws[lk + 1][lr] = wv.y; Ws[lk + 2][lr] = wv.z; Ws[lk + 3][lr] = wv.w;
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 w = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
      const float ar[4] = {a.x, a.y, a.z, a.w};
      const float wr[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], wr[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? bias[n] : 0.f);
      if (RELU) v = fmaxf(v, 0.f);
      C[(size_t)m * N + n] = v;
    }
  }
}

}  // namespace

int launch_linear_simt(const float* A1, int K1, const float* A2, int K2, const float* W, const float* bias, float* C,
                       int N, const int* m_dev, int m_max, bool relu, cudaStream_t stream) {
  MT_CHECK_ARG(K1 > 0 && K1 % BK == 0 && K2 >= 0 && K2 % BK == 0,
               "linear: input widths must be multiples of %d (got %d and %d)", BK, K1, K2);
  MT_CHECK_ARG(N > 0 && m_max >= 0, "linear: bad output shape (%d x %d)", m_max, N);
  if (m_max == 0) return MESHTOK_OK;
  dim3 grid(ceil_div(m_max, BM), ceil_div(N, BN));
  if (relu) linear_simt_kernel<true><<<grid, 256, 0, stream>>>(A1, K1, A2, K2, W, bias, C, N, m_dev, m_max);
  else linear_simt_kernel<false><<<grid, 256, 0, stream>>>(A1, K1, A2, K2, W, bias, C, N, m_dev, m_max);
  MT_CHECK_LAUNCH();
  return MESHTOK_OK;
}

}  // namespace meshtok
