// Token-embedding front end of MeshTransformer.forward_on_codes, meshgpt_pytorch.py:1526-1570:
//   codes (b, L) int64 -> fine tokens (b, ceil(L / n) * n, D) = token_embed[code] + abs_pos_emb[t] + level + vertex  (:1528-1547, :1560)
//                      -> face tokens (b, L / n, D) = LayerNorm(Linear(n * D -> D)(the n fine tokens of a face))     (:1564-1570)
// n = nvf * Q codes per face.  Both kernels are HBM-bound row gathers; neither runs a GEMM:
//   * fine_tokens_kernel adds the four rows in the reference's order ((tok + pos) + level) + vertex, so the result is bit-exact.
//   * face_tokens_kernel uses the Linear's linearity: W [x_0 .. x_{n-1}] = sum_j W_j (tok[c_j] + pos / level / vertex part), so with
//     the tables T_j = token_embed W_j^T (n tables of (K + 1, D), folded once at load time, frontend.py) and the per-face row
//     P[f] = sum_j W_j (pos[n f + j] + level + vertex) + bias, a face token is n + 1 row reads, a sum and a LayerNorm: 4 (n + 2) D
//     bytes of traffic per face instead of a (n D x D) product per face (3.1 MFLOP at n = 6, D = 512).
#include "common.cuh"
#include "kernels.cuh"

namespace meshtok {
namespace {

// A warp owns (position t, 128-column chunk, block of up to 32 meshes): the position / level / vertex rows of t are the same for
// every mesh, so they are read once into registers and only the gathered token_embed row and the output row move per token (the
// first version, a warp per token reading all four rows, ran at the L2's bandwidth: 3 row reads from L2 per row written).
// Tokens t >= L of the last (incomplete) face are zero rows (pad_to_length, :1560).
__global__ void __launch_bounds__(256) fine_tokens_kernel(const int64_t* __restrict__ codes, int B, int L, int Lp, int D, int64_t pad_id, int Q,
                                                          int nvf, int n_embed, const float* __restrict__ tok, const float* __restrict__ pos,
                                                          const float* __restrict__ level, const float* __restrict__ vert,
                                                          float* __restrict__ out) {
  const int lane = lane_id();
  const int chunks = (D + 127) >> 7, bblocks = (B + 31) >> 5;
  const long long total = (long long)Lp * chunks * bblocks;
  for (long long w = (long long)blockIdx.x * 8 + warp_id(); w < total; w += (long long)gridDim.x * 8) {
    const int bb = (int)(w % bblocks);
    const long long r = w / bblocks;
    const int ch = (int)(r % chunks), t = (int)(r / chunks);
    const int col = ch * 128 + lane * 4;
    const bool in_row = col < D;                                          // D % 4 == 0: a lane's float4 is inside the row or not at all
    const int b0 = bb * 32, nb = min(32, B - b0);
    if (t >= L) {
      if (in_row)
        for (int i = 0; i < nb; ++i) __stcs(reinterpret_cast<float4*>(out + ((size_t)(b0 + i) * Lp + t) * D + col), make_float4(0.f, 0.f, 0.f, 0.f));
      continue;
    }
    long long c = lane < nb ? codes[(size_t)(b0 + lane) * L + t] : 0;
    if (c == pad_id) c = 0;                                               // :1528
    c = c < 0 ? 0 : (c >= n_embed ? n_embed - 1 : c);                     // the reference's embedding would raise; stay in bounds
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f), l = p, v = p;
    if (in_row) {
      p = __ldg(reinterpret_cast<const float4*>(pos + (size_t)t * D + col));
      l = __ldg(reinterpret_cast<const float4*>(level + (size_t)(t % Q) * D + col));             // '(r q) d'        :1541
      v = __ldg(reinterpret_cast<const float4*>(vert + (size_t)((t / Q) % nvf) * D + col));      // '(r nv q) d'     :1546
    }
#pragma unroll 4
    for (int i = 0; i < nb; ++i) {
      const long long ci = __shfl_sync(0xffffffffu, c, i);
      if (!in_row) continue;
      const float4 a = __ldg(reinterpret_cast<const float4*>(tok + (size_t)ci * D + col));
      float4 o;
      o.x = __fadd_rn(__fadd_rn(__fadd_rn(a.x, p.x), l.x), v.x);
      o.y = __fadd_rn(__fadd_rn(__fadd_rn(a.y, p.y), l.y), v.y);
      o.z = __fadd_rn(__fadd_rn(__fadd_rn(a.z, p.z), l.z), v.z);
      o.w = __fadd_rn(__fadd_rn(__fadd_rn(a.w, p.w), l.w), v.w);
      __stcs(reinterpret_cast<float4*>(out + ((size_t)(b0 + i) * Lp + t) * D + col), o);   // written once: keep the tables in L2
    }
  }
}

// a warp per face; lane owns columns lane * 4 + 128 j (NJ = D / 128 float4 per lane); the n table rows are summed in slot order on
// top of the face's position row, then LayerNorm over the row (two-pass, in registers)
template <int NJ>
__global__ void __launch_bounds__(256) face_tokens_kernel(const int64_t* __restrict__ codes, int B, int L, int n, int64_t pad_id, int n_embed,
                                                          const float* __restrict__ tables, const float* __restrict__ face_pos,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                                          float* __restrict__ out) {
  constexpr int D = NJ * 128;
  const int nf = L / n, lane = lane_id();
  const long long total = (long long)B * nf;
  for (long long w = (long long)blockIdx.x * 8 + warp_id(); w < total; w += (long long)gridDim.x * 8) {
    const int b = (int)(w / nf), f = (int)(w - (long long)b * nf);
    long long c = lane < n ? codes[(size_t)b * L + (size_t)f * n + lane] : 0;
    if (c == pad_id) c = 0;
    c = c < 0 ? 0 : (c >= n_embed ? n_embed - 1 : c);
    float4 acc[NJ];
    const float4* pr = reinterpret_cast<const float4*>(face_pos + (size_t)f * D) + lane;
#pragma unroll
    for (int j = 0; j < NJ; ++j) acc[j] = __ldg(pr + 32 * j);
    for (int s = 0; s < n; ++s) {
      const long long cs = __shfl_sync(0xffffffffu, c, s);
      const float4* tr = reinterpret_cast<const float4*>(tables + ((size_t)s * n_embed + (size_t)cs) * D) + lane;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const float4 v = __ldg(tr + 32 * j);
        acc[j].x += v.x; acc[j].y += v.y; acc[j].z += v.z; acc[j].w += v.w;
      }
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) sum += (acc[j].x + acc[j].y) + (acc[j].z + acc[j].w);
    const float mean = warp_sum(sum) * (1.f / (float)D);
    float var = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const float dx = acc[j].x - mean, dy = acc[j].y - mean, dz = acc[j].z - mean, dw = acc[j].w - mean;
      var += (dx * dx + dy * dy) + (dz * dz + dw * dw);
    }
    const float rstd = rsqrtf(warp_sum(var) * (1.f / (float)D) + eps);
    float4* dst = reinterpret_cast<float4*>(out + (size_t)w * D) + lane;
    const float4* gr = reinterpret_cast<const float4*>(gamma) + lane;
    const float4* br = reinterpret_cast<const float4*>(beta) + lane;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const float4 g = __ldg(gr + 32 * j), bb = __ldg(br + 32 * j);
      float4 o;
      o.x = fmaf((acc[j].x - mean) * rstd, g.x, bb.x);
      o.y = fmaf((acc[j].y - mean) * rstd, g.y, bb.y);
      o.z = fmaf((acc[j].z - mean) * rstd, g.z, bb.z);
      o.w = fmaf((acc[j].w - mean) * rstd, g.w, bb.w);
      __stcs(dst + 32 * j, o);
    }
  }
}

inline unsigned warp_grid(long long warps) {
  const long long want = (warps + 7) / 8;
  return (unsigned)(want < 1 ? 1 : (want < 148 * 32 ? want : 148 * 32));
}

}  // namespace
}  // namespace meshtok

using namespace meshtok;

extern "C" {

int meshtok_frontend_fine_tokens(const int64_t* codes, int B, int L, int D, int64_t pad_id, int Q, int nvf, int n_embed,
                                 const float* token_embed, const float* abs_pos_emb, const float* level_embed, const float* vertex_embed,
                                 float* out, meshtok_stream_t stream) {
  MT_CHECK_ARG(codes && token_embed && abs_pos_emb && level_embed && vertex_embed && out && B >= 0 && L >= 0 && Q >= 1 && nvf >= 1 && n_embed >= 1,
               "frontend_fine_tokens: null pointer or bad size");
  MT_CHECK_ARG(D > 0 && D % 4 == 0, "frontend_fine_tokens: dim %d must be a multiple of 4", D);
  const int n = Q * nvf;
  const int Lp = (L + n - 1) / n * n;
  if (B == 0 || Lp == 0) return MESHTOK_OK;
  fine_tokens_kernel<<<warp_grid((long long)Lp * ((D + 127) / 128) * ((B + 31) / 32)), 256, 0, static_cast<cudaStream_t>(stream)>>>(codes, B, L, Lp, D, pad_id, Q, nvf, n_embed, token_embed,
                                                                                                  abs_pos_emb, level_embed, vertex_embed, out);
  MT_CHECK_LAUNCH();
  return MESHTOK_OK;
}

int meshtok_frontend_face_tokens(const int64_t* codes, int B, int L, int n, int D, int64_t pad_id, int n_embed, const float* folded_tables,
                                 const float* face_pos, const float* ln_gamma, const float* ln_beta, float eps, float* out,
                                 meshtok_stream_t stream) {
  MT_CHECK_ARG(n >= 1 && n <= 32, "frontend_face_tokens: %d codes per face (1 .. 32 supported)", n);
  MT_CHECK_ARG(D >= 128 && D <= 1024 && D % 128 == 0, "frontend_face_tokens: dim %d must be a multiple of 128, at most 1024", D);
  MT_CHECK_ARG(B >= 0 && L >= 0 && n_embed >= 1, "frontend_face_tokens: bad size");
  if (B == 0 || L / n == 0) return MESHTOK_OK;                           // no complete face: nothing to write (out may be NULL)
  MT_CHECK_ARG(codes && folded_tables && face_pos && ln_gamma && ln_beta && out, "frontend_face_tokens: null pointer");
  const unsigned grid = warp_grid((long long)B * (L / n));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
#define MT_FACE_TOKENS(NJ)                                                                                                                    \
  case NJ:                                                                                                                                    \
    face_tokens_kernel<NJ><<<grid, 256, 0, s>>>(codes, B, L, n, pad_id, n_embed, folded_tables, face_pos, ln_gamma, ln_beta, eps, out);     \
    break;
  switch (D / 128) {
    MT_FACE_TOKENS(1) MT_FACE_TOKENS(2) MT_FACE_TOKENS(3) MT_FACE_TOKENS(4) MT_FACE_TOKENS(5) MT_FACE_TOKENS(6) MT_FACE_TOKENS(7) MT_FACE_TOKENS(8)
  }
#undef MT_FACE_TOKENS
  MT_CHECK_LAUNCH();
  return MESHTOK_OK;
}

}  // extern "C"
