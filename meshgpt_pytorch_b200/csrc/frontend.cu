// Token-embedding front end of MeshTransformer.forward_on_codes, meshgpt_pytorch.py:1526-1570:
//   codes (b, L) int64 -> fine tokens (b, ceil(L / n) * n, D) = token_embed[code] + abs_pos_emb[t] + level + vertex  (:1528-1547, :1560)
//                      -> face tokens (b, L / n, D) = LayerNorm(Linear(n * D -> D)(the n fine tokens of a face))     (:1564-1570)
// n = nvf * Q codes per face.  Both kernels are HBM-bound row gathers; neither runs a GEMM:
//   * fine_tokens_kernel adds the four rows in the reference's order ((tok + pos) + level) + vertex, so the result is bit-exact.
//   * face_tokens_kernel uses the Linear's linearity: W [x_0 .. x_{n-1}] = sum_j W_j (tok[c_j] + pos / level / vertex part), so with
//     the tables T_j = token_embed W_j^T (n tables of (K + 1, D), folded once at load time, frontend.py) and the per-face row
//     P[f] = sum_j W_j (pos[n f + j] + level + vertex) + bias, a face token is n + 1 row reads, a sum and a LayerNorm: 4 (n + 2) D
//     bytes of traffic per face instead of a (n D x D) product per face (3.1 MFLOP at n = 6, D = 512).
#include <string.h>

#include "common.cuh"
#include "kernels.cuh"

namespace meshtok {
namespace {

// Where a token's code comes from: the (B, L) int64 tensor tokenize returns, or - skipping that tensor - the tokenizer's own result:
// the per-vertex codes (n_vrows, Q) int32 + the face / vertex maps of its workspace (token t = (face, slot, level) -> vertex id ->
// quantizer row -> code; pad_id on padded faces).  eos >= 0: the sequence is one token longer, the eos id sits at the first pad
// position of every row (append_eos, meshgpt_pytorch.py:1506-1518) and the extra last column is 0 unless the eos landed there.
struct CodeSource {
  const int64_t* codes;       // (B, L) or null
  meshtok_vertex_code_source v;
  int64_t pad_id, eos;        // eos < 0: no eos
  int L;                      // length of the UN-extended sequence (columns of `codes` / F * nvf * Q)
};
__device__ __forceinline__ long long raw_code(const CodeSource& s, int b, int t) {
  if (s.codes != nullptr) return s.codes[(size_t)b * s.L + t];
  const int tok = t / s.v.Q, q = t - tok * s.v.Q;
  const int f = tok / s.v.nvf;
  if (s.v.face_row[(size_t)b * s.v.F + f] < 0) return s.pad_id;
  const long long vid = s.v.faces[((size_t)b * s.v.F) * s.v.nvf + tok];
  return s.v.vertex_codes[(size_t)s.v.vert_row[(size_t)b * s.v.V + vid] * s.v.Q + q];
}
// code of position t of the (possibly eos-extended) sequence of mesh b; len_b = first pad position of the row (only needed with eos)
__device__ __forceinline__ long long seq_code(const CodeSource& s, int b, int t, int len_b) {
  if (s.eos >= 0) {
    if (t == len_b) return s.eos;
    if (t >= s.L) return 0;
  }
  return raw_code(s, b, t);
}
// first pad position of row b (the number of leading non-pad tokens, :1509); one warp, every lane gets the result
__device__ __forceinline__ int row_code_len(const CodeSource& s, int b) {
  const int lane = lane_id();
  for (int t0 = 0; t0 < s.L; t0 += 32) {
    const int t = t0 + lane;
    const bool pad = t < s.L && raw_code(s, b, t) == s.pad_id;
    const unsigned m = __ballot_sync(0xffffffffu, pad);
    if (m) return t0 + __ffs(m) - 1;
  }
  return s.L;
}

__global__ void __launch_bounds__(256) code_lens_kernel(CodeSource s, int B, int* __restrict__ lens) {
  const int b = blockIdx.x * 8 + warp_id();
  if (b >= B) return;
  const int len = row_code_len(s, b);
  if (lane_id() == 0) lens[b] = len;
}

// append_eos (:1506-1518) as a tensor: out (B, L + 1) int64
__global__ void __launch_bounds__(256) append_eos_kernel(CodeSource s, int B, const int* __restrict__ lens, int64_t* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int Le = s.L + 1;
  if (i >= (long long)B * Le) return;
  const int b = (int)(i / Le), t = (int)(i - (long long)b * Le);
  out[i] = seq_code(s, b, t, lens[b]);
}

// A warp owns (position t, 128-column chunk, block of up to 32 meshes): the position / level / vertex rows of t are the same for
// every mesh, so they are read once into registers and only the gathered token_embed row and the output row move per token (the
// first version, a warp per token reading all four rows, ran at the L2's bandwidth: 3 row reads from L2 per row written).
// Tokens t >= L of the last (incomplete) face are zero rows (pad_to_length, :1560).
__global__ void __launch_bounds__(256) fine_tokens_kernel(CodeSource src, const int* __restrict__ lens, int B, int L, int Lp, int D, int64_t pad_id, int Q,
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
    long long c = lane < nb ? seq_code(src, b0 + lane, t, lens ? lens[b0 + lane] : 0) : 0;
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
__global__ void __launch_bounds__(256) face_tokens_kernel(CodeSource src, const int* __restrict__ lens, int B, int L, int n, int64_t pad_id, int n_embed,
                                                          const float* __restrict__ tables, const float* __restrict__ face_pos,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                                          const float* __restrict__ sos, int n_sos, float* __restrict__ out) {
  constexpr int D = NJ * 128;
  const int nf = L / n, lane = lane_id();
  const int rows_out = nf + n_sos;   // sos rows first (auto prepend, :1591-1594), then the face tokens
  const long long total = (long long)B * rows_out;
  for (long long w = (long long)blockIdx.x * 8 + warp_id(); w < total; w += (long long)gridDim.x * 8) {
    const int b = (int)(w / rows_out), f = (int)(w - (long long)b * rows_out) - n_sos;
    if (f < 0) {   // a start-of-sequence row: the parameter itself
      const float4* sr = reinterpret_cast<const float4*>(sos + (size_t)(f + n_sos) * D) + lane;
      float4* dst = reinterpret_cast<float4*>(out + (size_t)w * D) + lane;
#pragma unroll
      for (int j = 0; j < NJ; ++j) __stcs(dst + 32 * j, __ldg(sr + 32 * j));
      continue;
    }
    long long c = lane < n ? seq_code(src, b, f * n + lane, lens ? lens[b] : 0) : 0;
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

static int make_source(const int64_t* codes, const meshtok_vertex_code_source* vs, int L, int64_t pad_id, int64_t eos, CodeSource* out) {
  memset(out, 0, sizeof(*out));
  MT_CHECK_ARG((codes != nullptr) != (vs != nullptr), "frontend: give either codes or a vertex-code source");
  out->codes = codes; out->pad_id = pad_id; out->eos = eos; out->L = L;
  if (vs != nullptr) {
    MT_CHECK_ARG(vs->faces && vs->face_row && vs->vert_row && vs->vertex_codes && vs->F > 0 && vs->nvf > 0 && vs->V > 0 && vs->Q > 0,
                 "frontend: incomplete vertex-code source");
    MT_CHECK_ARG(L == vs->F * vs->nvf * vs->Q, "frontend: L (%d) must be F * nvf * Q (%d) with a vertex-code source", L, vs->F * vs->nvf * vs->Q);
    out->v = *vs;
  }
  return MESHTOK_OK;
}

static int run_code_lens(const CodeSource& src, int B, int* lens_ws, cudaStream_t s) {
  MT_CHECK_ARG(lens_ws != nullptr, "frontend: append_eos needs a (B) int32 workspace for the row lengths");
  code_lens_kernel<<<(B + 7) / 8, 256, 0, s>>>(src, B, lens_ws);
  MT_CHECK_LAUNCH();
  return MESHTOK_OK;
}

int meshtok_frontend_append_eos(const int64_t* codes, const meshtok_vertex_code_source* vsrc, int B, int L, int64_t pad_id, int64_t eos_id,
                                int32_t* lens_ws, int64_t* out, meshtok_stream_t stream) {
  MT_CHECK_ARG(out != nullptr && B >= 0 && L >= 0 && eos_id >= 0, "frontend_append_eos: null pointer or bad size");
  CodeSource src;
  int rc = make_source(codes, vsrc, L, pad_id, eos_id, &src);
  if (rc != MESHTOK_OK) return rc;
  if (B == 0) return MESHTOK_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if ((rc = run_code_lens(src, B, lens_ws, s)) != MESHTOK_OK) return rc;
  const long long n = (long long)B * (L + 1);
  append_eos_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(src, B, lens_ws, out);
  MT_CHECK_LAUNCH();
  return MESHTOK_OK;
}

int meshtok_frontend_fine_tokens(const int64_t* codes, const meshtok_vertex_code_source* vsrc, int B, int L, int D, int64_t pad_id, int64_t eos_id,
                                 int32_t* lens_ws, int Q, int nvf, int n_embed, const float* token_embed, const float* abs_pos_emb,
                                 const float* level_embed, const float* vertex_embed, float* out, meshtok_stream_t stream) {
  MT_CHECK_ARG(token_embed && abs_pos_emb && level_embed && vertex_embed && out && B >= 0 && L >= 0 && Q >= 1 && nvf >= 1 && n_embed >= 1,
               "frontend_fine_tokens: null pointer or bad size");
  MT_CHECK_ARG(D > 0 && D % 4 == 0, "frontend_fine_tokens: dim %d must be a multiple of 4", D);
  CodeSource src;
  int rc = make_source(codes, vsrc, L, pad_id, eos_id, &src);
  if (rc != MESHTOK_OK) return rc;
  const int n = Q * nvf;
  const int Le = L + (eos_id >= 0 ? 1 : 0);          // the sequence the embeddings are taken of
  const int Lp = (Le + n - 1) / n * n;
  if (B == 0 || Lp == 0) return MESHTOK_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (eos_id >= 0 && (rc = run_code_lens(src, B, lens_ws, s)) != MESHTOK_OK) return rc;
  fine_tokens_kernel<<<warp_grid((long long)Lp * ((D + 127) / 128) * ((B + 31) / 32)), 256, 0, s>>>(src, eos_id >= 0 ? lens_ws : nullptr, B, Le, Lp, D, pad_id, Q, nvf,
                                                                                                  n_embed, token_embed, abs_pos_emb, level_embed, vertex_embed, out);
  MT_CHECK_LAUNCH();
  return MESHTOK_OK;
}

int meshtok_frontend_face_tokens(const int64_t* codes, const meshtok_vertex_code_source* vsrc, int B, int L, int n, int D, int64_t pad_id, int64_t eos_id,
                                 int32_t* lens_ws, int n_embed, const float* folded_tables, const float* face_pos, const float* ln_gamma,
                                 const float* ln_beta, float eps, const float* sos, int n_sos, float* out, meshtok_stream_t stream) {
  MT_CHECK_ARG(n >= 1 && n <= 32, "frontend_face_tokens: %d codes per face (1 .. 32 supported)", n);
  MT_CHECK_ARG(D >= 128 && D <= 1024 && D % 128 == 0, "frontend_face_tokens: dim %d must be a multiple of 128, at most 1024", D);
  MT_CHECK_ARG(B >= 0 && L >= 0 && n_embed >= 1 && n_sos >= 0 && (n_sos == 0 || sos != nullptr), "frontend_face_tokens: bad size");
  CodeSource src;
  int rc = make_source(codes, vsrc, L, pad_id, eos_id, &src);
  if (rc != MESHTOK_OK) return rc;
  const int Le = L + (eos_id >= 0 ? 1 : 0);
  if (B == 0 || Le / n + n_sos == 0) return MESHTOK_OK;                  // no complete face and no sos row: nothing to write (out may be NULL)
  MT_CHECK_ARG(folded_tables && face_pos && ln_gamma && ln_beta && out, "frontend_face_tokens: null pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (eos_id >= 0 && (rc = run_code_lens(src, B, lens_ws, s)) != MESHTOK_OK) return rc;
  const int* lens = eos_id >= 0 ? lens_ws : nullptr;
  const unsigned grid = warp_grid((long long)B * (Le / n + n_sos));
#define MT_FACE_TOKENS(NJ)                                                                                                                    \
  case NJ:                                                                                                                                    \
    face_tokens_kernel<NJ><<<grid, 256, 0, s>>>(src, lens, B, Le, n, pad_id, n_embed, folded_tables, face_pos, ln_gamma, ln_beta, eps, sos, n_sos, out);     \
    break;
  switch (D / 128) {
    MT_FACE_TOKENS(1) MT_FACE_TOKENS(2) MT_FACE_TOKENS(3) MT_FACE_TOKENS(4) MT_FACE_TOKENS(5) MT_FACE_TOKENS(6) MT_FACE_TOKENS(7) MT_FACE_TOKENS(8)
  }
#undef MT_FACE_TOKENS
  MT_CHECK_LAUNCH();
  return MESHTOK_OK;
}

}  // extern "C"
