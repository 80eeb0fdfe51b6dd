"""Host-side mirror of the tokenization half of meshgpt_pytorch.MeshAutoencoder.

Same constructor keywords, same state-dict key names, same method signatures and tensor contracts as the
reference (meshgpt_pytorch/meshgpt_pytorch.py:428-1019) for
    tokenize / forward(return_codes=True) / encode / quantize,
but every FLOP runs in libmeshtok's sm_100a CUDA kernels through the C ABI (include/meshtok.h).  PyTorch is
used for parameter storage, device memory and streams only.  There is no CPU path: inputs must be CUDA
tensors, and importing this module fails if the library has not been built.

Out of scope: the decoder (decode, decode_from_codes_to_faces, the reconstruction loss),
the optional encoder attention blocks (attn_encoder_depth must be 0) and training.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor, nn

from . import _lib
from ._lib import MeshtokError, Model, Taps, TokenizeArgs, check, lib, ptr, stream_ptr


# ---------------------------------------------------------------------------------------------------
# parameter containers with the reference's key names
# ---------------------------------------------------------------------------------------------------


class SAGEConv(nn.Module):
    """Parameters of torch_geometric.nn.SAGEConv(in, out, normalize=True, project=True): lin (in->in, bias),
    lin_l (in->out, bias), lin_r (in->out, no bias).  Compute lives in the CUDA kernels."""

    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.lin = nn.Linear(dim_in, dim_in, bias=True)
        self.lin_l = nn.Linear(dim_in, dim_out, bias=True)
        self.lin_r = nn.Linear(dim_in, dim_out, bias=False)


class ResidualLFQ(nn.Module):
    """Parameters of vector_quantize_pytorch.ResidualLFQ that tokenization touches: project_in
    (dim -> log2 K) and project_out (log2 K -> dim)."""

    def __init__(self, dim: int, num_quantizers: int, codebook_size: int):
        super().__init__()
        bits = int(round(math.log2(codebook_size)))
        assert 2 ** bits == codebook_size, "ResidualLFQ needs a power-of-two codebook_size"
        self.project_in = nn.Linear(dim, bits)
        self.project_out = nn.Linear(bits, dim)


class _EuclideanCodebook(nn.Module):
    def __init__(self, codebook_size: int, dim: int, init_std: float):
        super().__init__()
        self.register_buffer("embed", torch.randn(1, codebook_size, dim) * init_std)
        self.register_buffer("embed_avg", self.embed.clone())
        self.register_buffer("cluster_size", torch.zeros(1, codebook_size))
        self.register_buffer("initted", torch.tensor([True]))


class _VectorQuantize(nn.Module):
    def __init__(self, codebook: _EuclideanCodebook):
        super().__init__()
        self._codebook = codebook


class ResidualVQ(nn.Module):
    """Buffers of vector_quantize_pytorch.ResidualVQ(shared_codebook=True): layers.{l}._codebook.embed
    (1, K, D), aliased across levels.  The reference k-means-initialises the codebook on the first batch;
    here it is an explicit tensor (random normal * init_std until a state dict is loaded) with initted=True."""

    def __init__(self, dim: int, num_quantizers: int, codebook_size: int, init_std: float = 0.05):
        super().__init__()
        shared = _EuclideanCodebook(codebook_size, dim, init_std)
        self.layers = nn.ModuleList([_VectorQuantize(shared) for _ in range(num_quantizers)])

    @property
    def codebook(self) -> Tensor:
        return self.layers[0]._codebook.embed[0]


# ---------------------------------------------------------------------------------------------------


class MeshAutoencoder(nn.Module):
    def __init__(
        self,
        num_discrete_coors=128,
        coor_continuous_range: Tuple[float, float] = (-1.0, 1.0),
        dim_coor_embed=64,
        num_discrete_area=128,
        dim_area_embed=16,
        num_discrete_normals=128,
        dim_normal_embed=64,
        num_discrete_angle=128,
        dim_angle_embed=16,
        encoder_dims_through_depth: Tuple[int, ...] = (64, 128, 256, 256, 576),
        dim_codebook=192,
        num_quantizers=2,
        codebook_size=16384,
        use_residual_lfq=True,
        sageconv_kwargs: dict = dict(normalize=True, project=True),
        attn_encoder_depth=0,
        pad_id=-1,
        quads=False,
        rvq_codebook_init_std=0.05,
        **decoder_and_training_kwargs,   # accepted for signature compatibility, unused by tokenization
    ):
        super().__init__()
        if attn_encoder_depth != 0:
            raise NotImplementedError("attn_encoder_depth > 0 is out of scope for the B200 tokenization path")
        if dict(sageconv_kwargs) != dict(normalize=True, project=True):
            raise NotImplementedError("only SAGEConv(normalize=True, project=True) is implemented")
        if len(encoder_dims_through_depth) > _lib.MAX_SAGE_LAYERS:
            raise NotImplementedError(f"at most {_lib.MAX_SAGE_LAYERS} SAGE layers")
        self.num_vertices_per_face = 4 if quads else 3
        nvf = self.num_vertices_per_face
        self.num_discrete_coors = num_discrete_coors
        self.num_discrete_area = num_discrete_area
        self.num_discrete_normals = num_discrete_normals
        self.num_discrete_angle = num_discrete_angle
        self.coor_continuous_range = tuple(float(x) for x in coor_continuous_range)
        self.dim_codebook = dim_codebook
        self.num_quantizers = num_quantizers
        self.codebook_size = codebook_size
        self.use_residual_lfq = use_residual_lfq
        self.encoder_dims = tuple(encoder_dims_through_depth)
        self.pad_id = pad_id

        self.coor_embed = nn.Embedding(num_discrete_coors, dim_coor_embed)
        self.angle_embed = nn.Embedding(num_discrete_angle, dim_angle_embed)
        self.area_embed = nn.Embedding(num_discrete_area, dim_area_embed)
        self.normal_embed = nn.Embedding(num_discrete_normals, dim_normal_embed)
        init_dim = dim_coor_embed * 3 * nvf + dim_angle_embed * nvf + dim_normal_embed * 3 + dim_area_embed
        self.project_in = nn.Linear(init_dim, dim_codebook)

        first, *rest = encoder_dims_through_depth
        self.init_sage_conv = SAGEConv(dim_codebook, first)
        self.init_encoder_act_and_norm = nn.Sequential(nn.SiLU(), nn.LayerNorm(first))
        self.encoders = nn.ModuleList()
        cur = first
        for d in rest:
            self.encoders.append(SAGEConv(cur, d))
            cur = d
        self.project_dim_codebook = nn.Linear(cur, dim_codebook * nvf)
        if use_residual_lfq:
            self.quantizer = ResidualLFQ(dim_codebook, num_quantizers, codebook_size)
        else:
            self.quantizer = ResidualVQ(dim_codebook, num_quantizers, codebook_size, rvq_codebook_init_std)

        for p in self.parameters():
            p.requires_grad_(False)
        self._prep: Optional[dict] = None
        self._prep_key = None
        self._ws_cache: Dict[object, Tensor] = {}     # one grow-only workspace per tag (see _workspace)
        self._status_pin: Optional[Tensor] = None     # pinned ring of status words of calls whose check is deferred
        self._status_pending: list = []               # (event, ring slot, context) of those calls, oldest first
        self._status_next = 0
        self.force_gemm_simt = False      # parity tests: fp32 SIMT linears instead of tcgen05 split-bf16
        self.force_rvq_exact = False      # parity tests: fp32 SIMT nearest-code scan instead of tcgen05 screening
        self.keep_taps = False            # parity tests: also store every intermediate activation as fp32 rows (see taps())
        self.edge_slots_per_face = 8      # initial CSR capacity; grown automatically on overflow
        self.last_workspace: Optional[Tensor] = None
        self.last_shape: Optional[tuple] = None

    # ---- weights ----------------------------------------------------------------------------------
    def load_reference_state_dict(self, state_dict: Dict[str, Tensor]):
        """Loads a reference MeshAutoencoder.state_dict(); decoder / loss keys are ignored, every key of the
        tokenization half must be present."""
        own = self.state_dict()
        missing = [k for k in own if k not in state_dict]
        if missing:
            raise KeyError(f"state dict lacks tokenization keys: {missing[:8]}{'...' if len(missing) > 8 else ''}")
        self.load_state_dict({k: state_dict[k] for k in own}, strict=True)
        self._prep = None
        return [k for k in state_dict if k not in own]

    @property
    def device(self):
        return self.project_in.weight.device

    def _weights_key(self):
        return tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))

    def _sage_layers(self):
        return [self.init_sage_conv, *self.encoders]

    def _prepare(self) -> dict:
        key = self._weights_key()
        if self._prep is not None and self._prep_key == key:
            return self._prep
        dev = self.device
        if dev.type != "cuda":
            raise RuntimeError("MeshAutoencoder must live on a CUDA device (call .cuda()); there is no CPU path")
        keep = {}
        nvf, D = self.num_vertices_per_face, self.dim_codebook
        f32 = dict(dtype=torch.float32, device=dev)

        # fold Embedding -> project_in into one (bins, D) table per slot of the concat (meshgpt_pytorch.py:735-741)
        W = self.project_in.weight.detach().double()
        tables, col = [], 0
        for emb, reps in ((self.coor_embed, nvf * 3), (self.angle_embed, nvf), (self.area_embed, 1), (self.normal_embed, 3)):
            E = emb.weight.detach().double()
            d = E.shape[1]
            for _ in range(reps):
                tables.append(E @ W[:, col:col + d].t())
                col += d
        assert col == W.shape[1]
        keep["tables"] = torch.cat(tables, 0).to(**f32).contiguous()
        keep["pin_b"] = self.project_in.bias.detach().to(**f32).contiguous()

        m = Model()
        m.abi_version = _lib.ABI_VERSION
        m.nvf, m.dim_codebook, m.num_quantizers, m.codebook_size = nvf, D, self.num_quantizers, self.codebook_size
        m.use_lfq = int(self.use_residual_lfq)
        m.num_discrete_coors, m.num_discrete_angle = self.num_discrete_coors, self.num_discrete_angle
        m.num_discrete_area, m.num_discrete_normals = self.num_discrete_area, self.num_discrete_normals
        m.coor_lo, m.coor_hi = self.coor_continuous_range
        m.num_slots = nvf * 4 + 4
        m.folded_tables, m.project_in_bias = ptr(keep["tables"]), ptr(keep["pin_b"])

        def image_of(w: Tensor, name: str):
            n, k = w.shape
            nbytes = C.c_size_t()
            if lib.meshtok_linear_image_bytes(n, k, C.byref(nbytes)) != 0:
                return 0                       # shape not covered by the tcgen05 kernel -> SIMT for this model
            img = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
            check(lib.meshtok_linear_build_image(ptr(w), n, k, ptr(img), stream_ptr()))
            keep[name] = img
            return ptr(img)

        layers = self._sage_layers()
        m.num_sage_layers = len(layers)
        all_images = True
        with torch.cuda.device(dev):
            for i, conv in enumerate(layers):
                lw = conv.lin.weight.detach().to(**f32).contiguous()
                lb = conv.lin.bias.detach().to(**f32).contiguous()
                cw = torch.cat([conv.lin_l.weight.detach(), conv.lin_r.weight.detach()], dim=1).to(**f32).contiguous()
                cb = conv.lin_l.bias.detach().to(**f32).contiguous()
                keep[f"sage{i}"] = (lw, lb, cw, cb)
                s = m.sage[i]
                s.dim_in, s.dim_out = lw.shape[0], cw.shape[0]
                s.lin_w, s.lin_b, s.cat_w, s.lin_l_b = ptr(lw), ptr(lb), ptr(cw), ptr(cb)
                s.lin_w_img = image_of(lw, f"sage{i}_lin_img")
                s.cat_w_img = image_of(cw, f"sage{i}_cat_img")
                all_images &= bool(s.lin_w_img) and bool(s.cat_w_img)
            ln = self.init_encoder_act_and_norm[1]
            keep["ln"] = (ln.weight.detach().to(**f32).contiguous(), ln.bias.detach().to(**f32).contiguous())
            m.ln_gamma, m.ln_beta = ptr(keep["ln"][0]), ptr(keep["ln"][1])
            keep["pdc"] = (self.project_dim_codebook.weight.detach().to(**f32).contiguous(),
                           self.project_dim_codebook.bias.detach().to(**f32).contiguous())
            m.pdc_w, m.pdc_b = ptr(keep["pdc"][0]), ptr(keep["pdc"][1])
            m.pdc_w_img = image_of(keep["pdc"][0], "pdc_img")
            all_images &= bool(m.pdc_w_img)

            rvq_tc_ok = False
            if self.use_residual_lfq:
                q = self.quantizer
                keep["lfq"] = tuple(t.detach().to(**f32).contiguous() for t in
                                    (q.project_in.weight, q.project_in.bias, q.project_out.weight, q.project_out.bias))
                m.lfq_in_w, m.lfq_in_b, m.lfq_out_w, m.lfq_out_b = (ptr(t) for t in keep["lfq"])
            else:
                cb = self.quantizer.codebook.detach().to(**f32).contiguous()
                e2 = (cb * cb).sum(dim=-1).contiguous()
                keep["cb"] = (cb, e2)
                m.codebook, m.codebook_sqnorm = ptr(cb), ptr(e2)
                m.codebook_max_norm = float(e2.max().sqrt()) * (1 + 1e-6)
                amax = float(cb.abs().max())
                scale = 1.0 if not (amax > 0 and math.isfinite(amax)) else 2.0 ** (-math.frexp(amax)[1])
                m.codebook_image_scale = scale
                nbytes = C.c_size_t()
                K, Dd = cb.shape
                if K % 128 == 0 and Dd in (64, 128, 192):
                    check(lib.meshtok_rvq_codebook_image_bytes(K, Dd, C.byref(nbytes)))
                    img = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
                    check(lib.meshtok_rvq_build_codebook_image(ptr(cb), ptr(e2), K, Dd, scale, m.codebook_max_norm, ptr(img), stream_ptr()))
                    keep["cb_img"] = img
                    m.codebook_f16_tiles = ptr(img)
                    rvq_tc_ok = True
        self._prep = dict(model=m, keep=keep, tc_gemm=all_images, tc_rvq=rvq_tc_ok, device=dev)
        self._prep_key = key
        self._ws_cache.clear()
        return self._prep

    # ---- the pipeline call ------------------------------------------------------------------------
    def _workspace(self, prep, B, F, V, E, cap, tag=None, total_faces=0, total_vertices=0) -> Tensor:
        """One grow-only buffer per tag (callers that run concurrently - two graphs in flight - use different tags and so keep
        apart).  Batches from a collate function change F and V nearly every step; the C side only needs workspace_bytes >= its
        carve total, so the buffer is reallocated only when a call needs more than it holds.  A captured graph pins the buffer it
        was captured on (GraphedTokenizer keeps a reference), so growing the tag's buffer later never invalidates a graph."""
        nbytes = C.c_size_t()     # (total_faces / total_vertices: ragged layout - activation rows by the sum of the mesh sizes, not B x the largest)
        check(lib.meshtok_tokenize_workspace_bytes_ragged(C.byref(prep["model"]), B, F, V, E, cap, int(total_faces), int(total_vertices), C.byref(nbytes)))
        ws = self._ws_cache.get(tag)
        if ws is None or ws.numel() < nbytes.value or ws.device != prep["device"]:
            ws = None
            self._ws_cache.pop(tag, None)             # release the old buffer before asking for the larger one
            ws = torch.empty(nbytes.value, dtype=torch.uint8, device=prep["device"])
            self._ws_cache[tag] = ws
        return ws

    # ---- deferred data-error checks: no host sync on the default path ------------------------------------------
    _STATUS_RING = 16

    def _defer_status(self, ws: Tensor, context: str):
        """Copies the call's status word into a pinned ring slot behind the kernels (an async 4-byte D2H on the current stream) and
        remembers an event; _poll_status() looks at the finished ones when the NEXT call starts, check_last_status() waits for all."""
        if self._status_pin is None:
            self._status_pin = torch.zeros(self._STATUS_RING, dtype=torch.int32).pin_memory()
        if len(self._status_pending) >= self._STATUS_RING:
            self._poll_status(block=True)
        slot = self._status_next % self._STATUS_RING
        self._status_next += 1
        self._status_pin[slot: slot + 1].copy_(ws[:4].view(torch.int32), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._status_pending.append((ev, slot, context))

    def _poll_status(self, block: bool):
        while self._status_pending:
            ev, slot, context = self._status_pending[0]
            if block:
                ev.synchronize()
            elif not ev.query():
                return
            self._status_pending.pop(0)
            status = int(self._status_pin[slot])
            if status & _lib.WS_STATUS_EDGE_OVERFLOW and not (status & ~_lib.WS_STATUS_EDGE_OVERFLOW):
                self.edge_slots_per_face *= 4          # later calls get the larger buffer; the flagged batch must be re-run
                raise MeshtokError(f"edge buffer overflow in an earlier {context} call (its codes are incomplete): edge_slots_per_face "
                                   f"raised to {self.edge_slots_per_face}, tokenize that batch again (check_status=True re-runs by itself)")
            self._raise_on_status(status)

    def _run(self, *, vertices, faces, face_edges=None, encoded_in=None, want_codes=True, want_encoded=False,
             want_face_coords=False, want_quantized=False, histogram=None, check_status=True, workspace_tag=None):
        """check_status: True - read the device-side status word after the call (one host sync; re-runs with a larger edge buffer
        by itself); "deferred" - copy it to pinned memory behind the kernels and raise when the next call starts or at
        check_last_status() (no sync); False - the caller checks (graphs, pipelines)."""
        prep = self._prepare()
        dev = prep["device"]
        if check_status == "deferred":
            self._poll_status(block=False)
        if not faces.is_cuda:
            raise RuntimeError("inputs must be CUDA tensors (no CPU path)")
        nvf, D, Q = self.num_vertices_per_face, self.dim_codebook, self.num_quantizers
        B, F, nv = faces.shape
        assert nv == nvf, f"faces must have {nvf} vertices per face"          # meshgpt_pytorch.py:707-709
        faces64 = faces.to(torch.int64).contiguous()
        if vertices is not None:
            vertices = vertices.to(torch.float32).contiguous()
            V = vertices.shape[1]
        else:
            V = int(faces64.amax().item()) + 1                                     # quantize(): meshgpt_pytorch.py:800-801
            V = max(V, 1)
        E = 0
        if face_edges is not None:
            face_edges = face_edges.to(torch.int64).contiguous()
            if face_edges.shape[1] == 0:
                face_edges = torch.full((B, 1, 2), self.pad_id, dtype=torch.int64, device=dev)
            E = face_edges.shape[1]
        enc_dim = self.encoder_dims[-1]
        if encoded_in is not None:
            encoded_in = encoded_in.to(torch.float32).contiguous()
            assert encoded_in.shape == (B, F, enc_dim)

        out = {}
        if want_codes:
            out["codes"] = torch.empty((B, F * nvf, Q), dtype=torch.int64, device=dev)
        if want_encoded:
            out["encoded"] = torch.empty((B, F, enc_dim), dtype=torch.float32, device=dev)
        if want_face_coords:
            out["face_coords"] = torch.empty((B, F, nvf * 3), dtype=torch.int64, device=dev)
        if want_quantized:
            out["quantized"] = torch.empty((B, F, nvf * D), dtype=torch.float32, device=dev)

        cap = max(1024, self.edge_slots_per_face * B * F, (B * E if E else 0))
        with torch.cuda.device(dev):
            while True:
                ws = self._workspace(prep, B, F, V, E, cap, workspace_tag)
                a = TokenizeArgs()
                a.vertices, a.faces, a.face_edges, a.encoded_in = ptr(vertices), ptr(faces64), ptr(face_edges), ptr(encoded_in)
                a.B, a.F, a.V, a.E = B, F, V, E
                a.pad_id, a.edge_capacity = int(self.pad_id), cap
                a.codes_out = ptr(out.get("codes"))
                a.encoded_out = ptr(out.get("encoded"))
                a.face_coords_out = ptr(out.get("face_coords"))
                a.quantized_out = ptr(out.get("quantized"))
                a.histogram = ptr(histogram)
                a.stop_after_encode = int(not want_codes and not want_quantized)
                a.rvq_exact = int(self.force_rvq_exact or not prep["tc_rvq"])
                a.gemm_simt = int(self.force_gemm_simt or not prep["tc_gemm"])
                a.keep_taps = int(self.keep_taps)
                check(lib.meshtok_tokenize(C.byref(prep["model"]), C.byref(a), ptr(ws), ws.numel(), stream_ptr()))
                self.last_workspace, self.last_shape = ws, (B, F, V, E, cap)
                if not check_status:
                    break
                if check_status == "deferred":
                    self._defer_status(ws, "tokenize")
                    break
                status = int(ws[:4].view(torch.int32).item())
                if status & _lib.WS_STATUS_EDGE_OVERFLOW and face_edges is None:
                    cap *= 4                       # rare: very high-valence meshes; rerun with a larger edge buffer
                    self.edge_slots_per_face *= 4
                    if histogram is not None:
                        raise MeshtokError("edge buffer overflow while accumulating a histogram; raise edge_slots_per_face")
                    continue
                self._raise_on_status(status)
                break
        return out

    @staticmethod
    def _raise_on_status(status: int):
        if status & _lib.WS_STATUS_VERTEX_RANGE:
            raise IndexError("faces reference a vertex index outside [0, num_vertices) (the reference's gather would fail too)")
        if status & _lib.WS_STATUS_EDGE_RANGE:
            raise IndexError("face_edges reference a face outside the valid faces of their own mesh")
        if status & _lib.WS_STATUS_RAGGED_BOUNDS:
            raise IndexError("ragged input: a mesh has more faces / vertices than max_faces / max_vertices, or its offsets decrease")
        if status & _lib.WS_STATUS_EDGE_OVERFLOW:
            raise MeshtokError("edge buffer overflow")

    def check_last_status(self):
        """Raises for any data error of the calls made so far: waits for the deferred checks (the default of tokenize / forward /
        tokenize_ragged) and reads the last call's status word (callers that ran with check_status=False)."""
        self._poll_status(block=True)
        if self.last_workspace is not None:
            self._raise_on_status(int(self.last_workspace[:4].view(torch.int32).item()))

    def taps(self) -> Dict[str, Tensor]:
        """Views of the intermediate tensors of the last call (stage-wise parity tests).  The fp32 activations
        (x0, layer{i}) are only stored when the call ran with self.keep_taps = True."""
        prep = self._prepare()
        B, F, V, E, cap, tot_f, tot_v = (tuple(self.last_shape) + (0, 0))[:7]
        t = Taps()
        check(lib.meshtok_tokenize_taps_ragged(C.byref(prep["model"]), B, F, V, E, cap, tot_f, tot_v, C.byref(t)))
        ws = self.last_workspace

        def view(off, n, dtype):
            size = torch.empty((), dtype=dtype).element_size()
            return ws[off: off + n * size].view(dtype)

        counts = view(t.counts, 4, torch.int32).tolist()
        nf, nvr, ne = counts[0], counts[1], counts[2]
        D, Q = self.dim_codebook, self.num_quantizers
        res = dict(n_faces=nf, n_vrows=nvr, n_edges=ne,
                   face_row=view(t.face_row, B * F, torch.int32).view(B, F),
                   vert_row=view(t.vert_row, B * V, torch.int32).view(B, V),
                   indptr=view(t.indptr, nf + 1, torch.int32),
                   src=view(t.src, min(ne, cap), torch.int32),
                   x0=view(t.x0, nf * D, torch.float32).view(nf, D),
                   averaged=view(t.averaged, nvr * D, torch.float32).view(nvr, D),
                   vertex_codes=view(t.vertex_codes, nvr * Q, torch.int32).view(nvr, Q),
                   rvq_exact_rechecks=int(view(t.rvq_stats, 4, torch.int32)[1]))
        for i, d in enumerate(self.encoder_dims):
            res[f"layer{i}"] = view(t.layer_out[i], nf * d, torch.float32).view(nf, d)
        return res

    # ---- reference API ----------------------------------------------------------------------------
    @torch.no_grad()
    def encode(self, *, vertices, faces, face_edges, face_mask, face_edges_mask, return_face_coordinates=False):
        """meshgpt_pytorch.py:686-785.  -> (b, nf, dims[-1]) fp32 [, (b, nf, nvf*3) int64].
        Rows of PADDED faces: the embedding rows are zero like the reference's (:771-773); the discretised coordinates are 0 there,
        where the reference returns the coordinates of vertex 0 (its pad -> 0 fill, :711) - values every caller masks out."""
        faces = faces.masked_fill(~face_mask[:, :, None], self.pad_id)
        face_edges = face_edges.masked_fill(~face_edges_mask[:, :, None], self.pad_id)
        out = self._run(vertices=vertices, faces=faces, face_edges=face_edges, want_codes=False, want_encoded=True,
                        want_face_coords=return_face_coordinates)
        if not return_face_coordinates:
            return out["encoded"]
        return out["encoded"], out["face_coords"]

    @torch.no_grad()
    def quantize(self, *, faces, face_mask, face_embed, pad_id=None, rvq_sample_codebook_temp=1.0):
        """meshgpt_pytorch.py:787-870.  -> (face_embed_output (b, nf, nvf*dim), codes (b, nf*nvf, q) int64,
        commit_loss).  Eval semantics: no codebook sampling noise, commit loss is zero.
        face_embed_output on PADDED faces is unspecified: the reference gathers its pad-vertex row there (the quantiser applied to the
        mean of project_dim_codebook's bias chunks, :814-829); here it is quantize(zero row) for LFQ and zeros for RVQ.  Decode and the
        losses mask those rows (:924, :1061-1064); the codes of padded tokens are pad_id exactly like the reference's."""
        faces = faces.masked_fill(~face_mask[:, :, None], self.pad_id)
        out = self._run(vertices=None, faces=faces, encoded_in=face_embed, want_codes=True, want_quantized=True)
        commit_loss = torch.zeros(self.num_quantizers, device=faces.device)
        return out["quantized"], out["codes"], commit_loss

    @torch.no_grad()
    def forward(self, *, vertices, faces, face_edges=None, return_codes=False, histogram=None, check_status="deferred", **kwargs):
        """meshgpt_pytorch.py:978-1019 with return_codes=True (the tokenization branch).  The training /
        reconstruction branches are out of scope.  No host sync: data errors (a vertex id out of range, ...) raise when the next
        call starts or at check_last_status(); check_status=True checks before returning (one sync)."""
        if not return_codes:
            raise NotImplementedError("only forward(return_codes=True) (tokenization) is implemented on this path")
        if kwargs.get("return_recon_faces"):
            raise AssertionError("cannot return reconstructed faces when just returning raw codes")
        return self._run(vertices=vertices, faces=faces, face_edges=face_edges, histogram=histogram,
                         check_status=check_status)["codes"]

    @torch.no_grad()
    def tokenize(self, vertices, faces, face_edges=None, **kwargs):
        """meshgpt_pytorch.py:949-976.  vertices (b, nv, 3), faces (b, nf, nvf) padded with pad_id, optional
        face_edges (b, e, 2) -> codes (b, nf*nvf, num_quantizers) int64 (batch-less inputs allowed)."""
        assert "return_codes" not in kwargs
        given = [t for t in (vertices, faces, face_edges) if t is not None]
        ndims = {t.ndim for t in given}
        assert len(ndims) == 1
        batch_less = ndims.pop() == 2
        if batch_less:
            vertices, faces = vertices[None], faces[None]
            face_edges = None if face_edges is None else face_edges[None]
        self.eval()
        codes = self.forward(vertices=vertices, faces=faces, face_edges=face_edges, return_codes=True, **kwargs)
        return codes[0] if batch_less else codes

    # ---- ragged input: a CSR of meshes instead of pad_id padding --------------------
    @torch.no_grad()
    def tokenize_ragged(self, vertices: Tensor, faces: Tensor, face_offsets, vertex_offsets, *, face_edges: Optional[Tensor] = None,
                        edge_offsets=None, max_faces: Optional[int] = None, max_vertices: Optional[int] = None,
                        histogram: Optional[Tensor] = None, return_face_coordinates: bool = False, check_status="deferred"):
        """tokenize() for a batch that was never padded: what the reference's custom_collate (data.py:347-369) turns into
        (b, nf, nvf) / (b, nv, 3) tensors full of pad_id stays a CSR of meshes here.

        vertices (sum nv, 3) float and faces (sum nf, nvf) integer with MESH-LOCAL vertex ids, both CUDA; face_offsets / vertex_offsets:
        B + 1 non-decreasing integers starting at 0 (list, CPU tensor or CUDA tensor): mesh b owns faces[face_offsets[b]:face_offsets[b+1]]
        and vertices[vertex_offsets[b]:vertex_offsets[b+1]].  Returns codes (sum nf * nvf, num_quantizers) int64, the token sequences of
        the meshes back to back: exactly tokenize(padded batch) with the pad tokens removed (same kernels, same summation orders).
        face_edges (sum ne, 2) + edge_offsets (B + 1), optional: precomputed edges (the face-edge cache's content without its padding)
        with MESH-LOCAL face indices, any order, duplicates allowed - the ragged form of tokenize(face_edges=...).
        max_faces / max_vertices: upper bounds of the per-mesh counts; computed from the offsets when omitted (a device sync if the
        offsets live on the GPU).  The same kernels run as for the padded layout; only the three that touch the inputs and the output
        (topology_kernel, face_bins_kernel, gather_codes_ragged_kernel) address them through the offsets."""
        r = self._ragged_args(vertices, faces, face_offsets, vertex_offsets, face_edges, edge_offsets, max_faces, max_vertices)
        nvf, Q = self.num_vertices_per_face, self.num_quantizers
        total, dev = r["total"], r["dev"]
        codes = torch.empty((total * nvf, Q), dtype=torch.int64, device=dev)
        coords = torch.empty((total, nvf * 3), dtype=torch.int64, device=dev) if return_face_coordinates else None
        if r["B"] == 0 or total == 0:
            return (codes, coords) if return_face_coordinates else codes
        self._launch_ragged(r["vertices"], r["faces"], r["fo"], r["vo"], r["B"], r["F"], r["V"], total, codes, coords, histogram, r["face_edges"], r["eo"],
                            r["n_edges"], check_status, "ragged")
        return (codes, coords) if return_face_coordinates else codes

    def _ragged_args(self, vertices, faces, face_offsets, vertex_offsets, face_edges=None, edge_offsets=None, max_faces=None, max_vertices=None):
        """Checks and device-side form of a ragged batch (shared by tokenize_ragged / encode_ragged / quantize_ragged)."""
        prep = self._prepare()
        dev = prep["device"]
        if not faces.is_cuda or (vertices is not None and not vertices.is_cuda):
            raise RuntimeError("inputs must be CUDA tensors (no CPU path)")
        nvf = self.num_vertices_per_face
        assert faces.ndim == 2 and faces.shape[1] == nvf, f"faces must be (sum nf, {nvf})"
        assert vertices is None or (vertices.ndim == 2 and vertices.shape[1] == 3), "vertices must be (sum nv, 3)"
        if self.training:
            self.eval()

        def offsets(o, total, what):
            o = torch.as_tensor(o)
            assert o.ndim == 1 and o.numel() >= 1 and not o.is_floating_point(), f"{what} must be B + 1 integers"
            if not o.is_cuda:     # host-side offsets (the usual case: the loader knows the mesh sizes) are checked without a device sync
                d = o[1:] - o[:-1]
                assert int(o[0]) == 0 and (total is None or int(o[-1]) == total) and bool((d >= 0).all()), f"{what} must rise from 0 to {total}"
                return o.to(device=dev, dtype=torch.int32), (int(d.max()) if d.numel() else 0)
            return o.to(torch.int32).contiguous(), None
        fo, fmax = offsets(face_offsets, faces.shape[0], "face_offsets")
        vo, vmax = offsets(vertex_offsets, None if vertices is None else vertices.shape[0], "vertex_offsets")
        assert fo.numel() == vo.numel(), "face_offsets and vertex_offsets must have B + 1 entries each"
        B = fo.numel() - 1
        eo, n_edges = None, 0
        if face_edges is not None:
            assert edge_offsets is not None and face_edges.is_cuda and face_edges.ndim == 2 and face_edges.shape[1] == 2, \
                "face_edges must be a CUDA tensor (sum ne, 2) with edge_offsets"
            n_edges = face_edges.shape[0]
            eo, _ = offsets(edge_offsets, n_edges, "edge_offsets")
            assert eo.numel() == B + 1, "edge_offsets must have B + 1 entries"
            face_edges = face_edges.to(torch.int64).contiguous()
            if n_edges == 0:      # an empty list still means "edges are given": one all-pad edge, like the padded layout
                face_edges = torch.full((1, 2), self.pad_id, dtype=torch.int64, device=dev)
        if max_faces is None:
            max_faces = fmax if fmax is not None else (int((fo[1:] - fo[:-1]).max()) if B else 0)
        if max_vertices is None:
            max_vertices = vmax if vmax is not None else (int((vo[1:] - vo[:-1]).max()) if B else 0)
        return dict(dev=dev, B=B, F=max(int(max_faces), 1), V=max(int(max_vertices), 1), total=faces.shape[0], fo=fo, vo=vo, eo=eo, n_edges=n_edges,
                    faces=faces.to(torch.int64).contiguous(), vertices=None if vertices is None else vertices.to(torch.float32).contiguous(),
                    face_edges=face_edges)

    @torch.no_grad()
    def encode_ragged(self, vertices: Tensor, faces: Tensor, face_offsets, vertex_offsets, *, face_edges: Optional[Tensor] = None,
                      edge_offsets=None, return_face_coordinates: bool = False):
        """encode() (meshgpt_pytorch.py:686-785) on a ragged batch: -> face embeddings (sum nf, dims[-1]) fp32, one row per face in input
        order (zero rows for faces that contain pad_id) [, discretised coordinates (sum nf, nvf*3) int64]."""
        r = self._ragged_args(vertices, faces, face_offsets, vertex_offsets, face_edges, edge_offsets)
        nvf, total, dev = self.num_vertices_per_face, r["total"], r["dev"]
        enc = torch.zeros((total, self.encoder_dims[-1]), dtype=torch.float32, device=dev)
        coords = torch.zeros((total, nvf * 3), dtype=torch.int64, device=dev) if return_face_coordinates else None
        if r["B"] and total:
            self._launch_ragged(r["vertices"], r["faces"], r["fo"], r["vo"], r["B"], r["F"], r["V"], total, None, coords, None, r["face_edges"], r["eo"],
                                r["n_edges"], True, "ragged", encoded_out=enc)
        return (enc, coords) if return_face_coordinates else enc

    @torch.no_grad()
    def quantize_ragged(self, faces: Tensor, face_offsets, vertex_offsets, face_embed: Tensor):
        """quantize() (meshgpt_pytorch.py:787-870) on a ragged batch: face_embed (sum nf, dims[-1]) -> (face_embed_output (sum nf, nvf*dim),
        codes (sum nf * nvf, q) int64, commit_loss).  vertex_offsets only carries the per-mesh vertex counts (ids are range-checked)."""
        r = self._ragged_args(None, faces, face_offsets, vertex_offsets)
        nvf, total, dev = self.num_vertices_per_face, r["total"], r["dev"]
        assert face_embed.is_cuda and tuple(face_embed.shape) == (total, self.encoder_dims[-1])
        codes = torch.empty((total * nvf, self.num_quantizers), dtype=torch.int64, device=dev)
        quant = torch.zeros((total, nvf * self.dim_codebook), dtype=torch.float32, device=dev)
        if r["B"] and total:
            self._launch_ragged(None, r["faces"], r["fo"], r["vo"], r["B"], r["F"], r["V"], total, codes, None, None, None, None, 0, True, "ragged",
                                encoded_in=face_embed.to(torch.float32).contiguous(), quantized_out=quant)
        return quant, codes, torch.zeros(self.num_quantizers, device=dev)

    def _launch_ragged(self, vertices, faces64, fo, vo, B, F, V, total, codes, coords=None, histogram=None, face_edges=None, eo=None,
                       n_edges=0, check_status=True, workspace_tag="ragged", encoded_in=None, encoded_out=None, quantized_out=None):
        """One meshtok_tokenize call in the ragged layout on prepared device tensors.  `total` is the number of face rows the buffers
        hold (the kernels take the live count from face_offsets[B] on the device, so it may be a capacity)."""
        prep = self._prepare()
        if check_status == "deferred":
            self._poll_status(block=False)
        E = 0 if face_edges is None else 1      # only the "edges are given" flag in this layout
        total_v = 0 if vertices is None else int(vertices.shape[0])      # 0: unknown - vertex rows reserved for B * V
        cap = max(1024, self.edge_slots_per_face * min(B * F, max(int(total), 1)), n_edges)
        with torch.cuda.device(prep["device"]):
            while True:
                ws = self._workspace(prep, B, F, V, E, cap, workspace_tag, total, total_v)
                a = TokenizeArgs()
                a.vertices, a.faces, a.face_offsets, a.vertex_offsets, a.total_faces = ptr(vertices), ptr(faces64), ptr(fo), ptr(vo), total
                a.total_vertices = total_v
                a.face_edges, a.edge_offsets, a.total_edges = ptr(face_edges), ptr(eo), n_edges
                a.B, a.F, a.V, a.E = B, F, V, E
                a.pad_id, a.edge_capacity = int(self.pad_id), cap
                a.codes_out, a.face_coords_out, a.histogram = ptr(codes), ptr(coords), ptr(histogram)
                a.encoded_in, a.encoded_out, a.quantized_out = ptr(encoded_in), ptr(encoded_out), ptr(quantized_out)
                a.stop_after_encode = int(codes is None and quantized_out is None)
                a.rvq_exact = int(self.force_rvq_exact or not prep["tc_rvq"])
                a.gemm_simt = int(self.force_gemm_simt or not prep["tc_gemm"])
                check(lib.meshtok_tokenize(C.byref(prep["model"]), C.byref(a), ptr(ws), ws.numel(), stream_ptr()))
                self.last_workspace, self.last_shape = ws, (B, F, V, E, cap, int(total), total_v)
                if not check_status:
                    break
                if check_status == "deferred":
                    self._defer_status(ws, "tokenize_ragged")
                    break
                status = int(ws[:4].view(torch.int32).item())
                if status & _lib.WS_STATUS_EDGE_OVERFLOW and histogram is None and face_edges is None:
                    cap *= 4
                    self.edge_slots_per_face *= 4
                    continue
                self._raise_on_status(status)
                break

    def graphed_ragged(self, max_meshes: int, max_faces: int, max_vertices: int, max_total_faces: int, max_total_vertices: int,
                       histogram: Optional[Tensor] = None) -> "GraphedRaggedTokenizer":
        """tokenize_ragged as ONE CUDA graph for every batch within the stated bounds (see GraphedRaggedTokenizer)."""
        return GraphedRaggedTokenizer(self, max_meshes, max_faces, max_vertices, max_total_faces, max_total_vertices, histogram)

    # ---- CUDA-graph replay of the whole pipeline ----------------------------------------------------
    def graphed(self, vertices: Tensor, faces: Tensor, histogram: Optional[Tensor] = None,
                face_edges: Optional[Tensor] = None) -> "GraphedTokenizer":
        """Captures tokenize() for this input shape into a CUDA graph (the C ABI never allocates or synchronises, so
        the ~40 kernel launches of a batch replay as one graph launch).  The example inputs are validated by one eager
        run first (status word, edge-buffer size).  See GraphedTokenizer."""
        return GraphedTokenizer(self, vertices, faces, histogram, face_edges)

    def host_pipeline(self, vertices: Tensor, faces: Tensor, depth: int = 2, histogram: Optional[Tensor] = None) -> "HostPipeline":
        """Multi-buffered host-to-host tokenization for a stream of equally shaped batches (see HostPipeline)."""
        return HostPipeline(self, vertices, faces, depth, histogram)

    def host_pipeline_ragged(self, max_meshes: int, max_faces: int, max_vertices: int, max_total_faces: int, max_total_vertices: int,
                             depth: int = 2, histogram: Optional[Tensor] = None) -> "RaggedHostPipeline":
        """The same ring for the ragged layout: one graph per slot serves batches of any composition (see RaggedHostPipeline)."""
        return RaggedHostPipeline(self, max_meshes, max_faces, max_vertices, max_total_faces, max_total_vertices, depth, histogram)

    # ---- host-buffer entry (what a CPU-side caller of the reference would do) ---------------------
    @torch.no_grad()
    def tokenize_host(self, vertices: Tensor, faces: Tensor, face_edges: Optional[Tensor] = None, pinned_out: Optional[Tensor] = None):
        """Host tensors in, host tensor out: H2D copies, the device pipeline and the D2H copy of the codes,
        all on the current stream; returns after the codes have landed."""
        dev = self.device
        v = vertices.to(dev, non_blocking=True)
        f = faces.to(dev, non_blocking=True)
        e = None if face_edges is None else face_edges.to(dev, non_blocking=True)
        codes = self.forward(vertices=v, faces=f, face_edges=e, return_codes=True, check_status=False)
        if pinned_out is None:
            pinned_out = torch.empty(codes.shape, dtype=codes.dtype, pin_memory=True)
        pinned_out.copy_(codes, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        self.check_last_status()
        return pinned_out


class GraphedTokenizer:
    """tokenize() for one input shape as a CUDA graph.

        g = model.graphed(vertices, faces)            # CUDA tensors of the shape to be replayed
        codes = g(vertices2, faces2)                  # copies into the static inputs, replays, returns g.codes
        codes = g()                                   # replays on whatever g.vertices / g.faces hold now

    g.codes is a static buffer that the next replay overwrites; clone it to keep it.  Data errors (vertex id out of
    range, edge-buffer overflow) are reported by g.check() / model.check_last_status(), not by the replay itself."""

    def __init__(self, model: MeshAutoencoder, vertices: Tensor, faces: Tensor, histogram: Optional[Tensor] = None,
                 face_edges: Optional[Tensor] = None):
        assert vertices.is_cuda and faces.is_cuda and vertices.ndim == 3 and faces.ndim == 3
        self.model = model
        # optional user-supplied (b, e, 2) edge lists (the reference's precomputed face_edges), a static input like the others
        self.face_edges = None if face_edges is None else face_edges.to(torch.int64).clone()
        tag = ("graph", id(self))             # its own workspace: several graphs of one model may be in flight
        self.vertices = vertices.to(torch.float32).clone()
        self.faces = faces.to(torch.int64).clone()
        self.histogram = histogram
        model.eval()
        # eager run: raises on bad data, grows the edge buffer if needed, allocates the workspace
        model._run(vertices=self.vertices, faces=self.faces, face_edges=self.face_edges, check_status=True, workspace_tag=tag)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            model._run(vertices=self.vertices, faces=self.faces, face_edges=self.face_edges, check_status=False, workspace_tag=tag)
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.codes = model._run(vertices=self.vertices, faces=self.faces, face_edges=self.face_edges, histogram=histogram,
                                    check_status=False, workspace_tag=tag)["codes"]
        self._workspace = model.last_workspace
        model._ws_cache.pop(tag, None)        # the graph owns its workspace from here on (no entry left behind when it is dropped)

    def __call__(self, vertices: Optional[Tensor] = None, faces: Optional[Tensor] = None, face_edges: Optional[Tensor] = None) -> Tensor:
        if vertices is not None:
            self.vertices.copy_(vertices, non_blocking=True)
        if faces is not None:
            self.faces.copy_(faces, non_blocking=True)
        if face_edges is not None:
            assert self.face_edges is not None and face_edges.shape == self.face_edges.shape, "captured without / with other face_edges"
            self.face_edges.copy_(face_edges, non_blocking=True)
        self.graph.replay()
        return self.codes

    def check(self):
        MeshAutoencoder._raise_on_status(int(self._workspace[:4].view(torch.int32).item()))

    def tokenize_host(self, vertices: Tensor, faces: Tensor, pinned_out: Optional[Tensor] = None) -> Tensor:
        """Host tensors in, host tensor out (H2D into the static inputs, one graph launch, D2H of the codes)."""
        codes = self(vertices, faces)
        if pinned_out is None:
            pinned_out = torch.empty(codes.shape, dtype=codes.dtype, pin_memory=True)
        pinned_out.copy_(codes, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        self.check()
        return pinned_out


class GraphedRaggedTokenizer:
    """tokenize_ragged() as a CUDA graph that serves EVERY batch within its bounds: at most max_meshes meshes of at most max_faces faces /
    max_vertices vertices each, max_total_faces face rows and max_total_vertices vertex rows in all.  The graph is captured once on
    static buffers of those sizes; the live sizes travel in the offset arrays (a batch with fewer meshes repeats its last offset: empty
    meshes), and the kernels take the live face count from face_offsets[B] on the device, so one graph covers ragged batches of any
    composition - the padded layout needs one graph per (b, nf, nv) shape.

        g = model.graphed_ragged(256, 800, 441, 256 * 500, 256 * 300)
        codes = g(vertices, faces, face_offsets, vertex_offsets)      # CUDA or (pinned) host tensors; -> (sum nf * nvf, Q) view of g.codes

    The returned view is overwritten by the next call; data errors are reported by g.check()."""

    def __init__(self, model: MeshAutoencoder, max_meshes: int, max_faces: int, max_vertices: int, max_total_faces: int,
                 max_total_vertices: int, histogram: Optional[Tensor] = None):
        prep = model._prepare()
        dev = prep["device"]
        self.model, self.nvf = model, model.num_vertices_per_face
        self.B, self.F, self.V = int(max_meshes), max(int(max_faces), 1), max(int(max_vertices), 1)
        self.cap_faces = min(int(max_total_faces), self.B * self.F)
        self.cap_vertices = min(int(max_total_vertices), self.B * self.V)
        assert self.B > 0 and self.cap_faces > 0 and self.cap_vertices > 0
        if model.training:
            model.eval()
        with torch.cuda.device(dev):
            self.vertices = torch.zeros((self.cap_vertices, 3), dtype=torch.float32, device=dev)
            self.faces = torch.zeros((self.cap_faces, self.nvf), dtype=torch.int64, device=dev)
            self.codes = torch.empty((self.cap_faces * self.nvf, model.num_quantizers), dtype=torch.int64, device=dev)
            self._stage = torch.empty(2 * (self.B + 1), dtype=torch.int32).pin_memory()    # host offsets -> device in one copy
            self._offsets_dev = torch.zeros(2 * (self.B + 1), dtype=torch.int32, device=dev)
            self.face_offsets, self.vertex_offsets = self._offsets_dev[: self.B + 1], self._offsets_dev[self.B + 1:]
            # the input the graph is captured on: one triangle (or quad) in mesh 0, every other mesh empty
            assert self.cap_vertices >= self.nvf and self.V >= self.nvf
            self.vertices[: self.nvf] = torch.eye(4, 3, device=dev)[: self.nvf] * 0.5
            self.faces[0] = torch.arange(self.nvf, device=dev)
            self.face_offsets[1:] = 1
            self.vertex_offsets[1:] = self.nvf
            tag = ("ragged-graph", id(self))
            args = (self.vertices, self.faces, self.face_offsets, self.vertex_offsets, self.B, self.F, self.V, self.cap_faces, self.codes)
            model._launch_ragged(*args, histogram=None, check_status=True, workspace_tag=tag)          # allocates the workspace
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                model._launch_ragged(*args, histogram=None, check_status=False, workspace_tag=tag)
            torch.cuda.current_stream().wait_stream(side)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                model._launch_ragged(*args, histogram=histogram, check_status=False, workspace_tag=tag)
        self._workspace = model.last_workspace
        model._ws_cache.pop(tag, None)
        self._stage_free = torch.cuda.Event()

    def __call__(self, vertices: Tensor, faces: Tensor, face_offsets, vertex_offsets) -> Tensor:
        fo, vo = torch.as_tensor(face_offsets), torch.as_tensor(vertex_offsets)
        n, nf, nv = fo.numel() - 1, faces.shape[0], vertices.shape[0]
        assert 0 <= n <= self.B and vo.numel() == n + 1, f"at most {self.B} meshes per call, B + 1 offsets each"
        assert nf <= self.cap_faces and nv <= self.cap_vertices, f"batch exceeds the captured bounds ({self.cap_faces} faces, {self.cap_vertices} vertices)"
        self.vertices[:nv].copy_(vertices, non_blocking=True)
        self.faces[:nf].copy_(faces, non_blocking=True)
        if fo.is_cuda:
            self.face_offsets[: n + 1].copy_(fo)
            self.vertex_offsets[: n + 1].copy_(vo)
            self.face_offsets[n + 1:] = self.face_offsets[n]
            self.vertex_offsets[n + 1:] = self.vertex_offsets[n]
        else:
            st = self._stage
            self._stage_free.synchronize()          # the previous call's copy out of the pinned staging buffer has finished
            st[: n + 1] = fo
            st[n + 1: self.B + 1] = int(fo[-1])
            st[self.B + 1: self.B + 2 + n] = vo
            st[self.B + 2 + n:] = int(vo[-1])
            self._offsets_dev.copy_(st, non_blocking=True)
            self._stage_free.record()
        self.graph.replay()
        return self.codes[: nf * self.nvf]

    def check(self):
        MeshAutoencoder._raise_on_status(int(self._workspace[:4].view(torch.int32).item()))


class HostPipeline:
    """Host tensors in, host tensors out, with the copies of one batch overlapping the kernels of the others - the native batch
    pipeline of libmeshtok (meshtok_pipeline_*, csrc/pipeline.cu): `depth` slots, each with its own stream, device staging, workspace,
    CUDA graph of the whole tokenize pipeline and pinned result buffer.  A submit is ONE C call (two H2D copies, a graph launch, the
    D2H of the codes and of the status word, an event); no Python-side stream juggling, no synchronising status read.

        pipe = model.host_pipeline(example_vertices, example_faces)      # CUDA tensors of the batch shape (a valid batch)
        for v, f in batches:                                             # pinned host tensors of that shape
            done = pipe.submit(v, f)        # -> (ticket, codes) of the batch submitted `depth` calls ago, or None
        for ticket, codes in pipe.drain(): ...

    The returned codes tensor is the slot's pinned buffer: copy it before the slot is used again (`depth` submits later).  A data
    error in a batch (vertex id out of range, ...) raises when that batch is handed back.  depth > 2 lets more batches overlap on
    the SMs at the price of one workspace per slot.  histogram (optional, (Q, K) int64 CUDA): accumulated by every batch."""

    def __init__(self, model: MeshAutoencoder, vertices: Tensor, faces: Tensor, depth: int = 2, histogram: Optional[Tensor] = None):
        assert depth >= 1 and vertices.is_cuda and faces.is_cuda and vertices.ndim == 3 and faces.ndim == 3
        prep = model._prepare()
        dev = prep["device"]
        model.eval()
        v0 = vertices.to(torch.float32).contiguous()
        f0 = faces.to(torch.int64).contiguous()
        B, F, nvf = f0.shape
        V = v0.shape[1]
        assert nvf == model.num_vertices_per_face
        model._run(vertices=v0, faces=f0, check_status=True)                 # raises on bad example data; sizes the edge buffer
        cap = max(1024, model.edge_slots_per_face * B * F)
        nbytes = C.c_size_t()
        check(lib.meshtok_tokenize_workspace_bytes(C.byref(prep["model"]), B, F, V, 0, cap, C.byref(nbytes)))
        self.model, self.depth, self.shape = model, depth, (B, F, V)
        self._keep = []
        slots = (_lib.PipelineSlot * depth)()
        self.outs, self._status = [], []
        with torch.cuda.device(dev):
            for s in range(depth):
                v, f = v0.clone(), f0.clone()
                codes = torch.empty((B, F * nvf, model.num_quantizers), dtype=torch.int64, device=dev)
                ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
                out = torch.empty(codes.shape, dtype=torch.int64).pin_memory()
                st = torch.zeros(4, dtype=torch.int32).pin_memory()
                self._keep += [v, f, codes, ws]
                self.outs.append(out)
                self._status.append(st)
                slots[s].vertices, slots[s].faces, slots[s].codes = ptr(v), ptr(f), ptr(codes)
                slots[s].workspace, slots[s].workspace_bytes = ptr(ws), ws.numel()
                slots[s].codes_host, slots[s].status_host = ptr(out), ptr(st)
            torch.cuda.synchronize(dev)
            handle = C.c_void_p()
            check(lib.meshtok_pipeline_create(C.byref(prep["model"]), B, F, V, int(model.pad_id), cap, ptr(histogram), slots, depth, C.byref(handle)))
        self._h = handle
        self._histogram = histogram
        self._res = _lib.PipelineResult()
        self.n = 0

    def streams(self):
        """The slots' CUDA streams as torch objects (to order events of the caller against the pipeline, e.g. for timing)."""
        return [torch.cuda.ExternalStream(lib.meshtok_pipeline_stream(self._h, s)) for s in range(self.depth)]

    def _result(self, r):
        if r.ticket < 0:
            return None
        if r.status:
            MeshAutoencoder._raise_on_status(int(r.status))
        return int(r.ticket), self.outs[r.slot]

    def submit(self, vertices: Tensor, faces: Tensor):
        B, F, V = self.shape
        assert not vertices.is_cuda and not faces.is_cuda, "HostPipeline takes HOST tensors (pinned for asynchronous copies)"
        assert vertices.dtype == torch.float32 and faces.dtype == torch.int64 and vertices.is_contiguous() and faces.is_contiguous()
        assert tuple(vertices.shape) == (B, V, 3) and tuple(faces.shape[:2]) == (B, F), "batch shape differs from the one the pipeline was built for"
        check(lib.meshtok_pipeline_submit(self._h, ptr(vertices), ptr(faces), C.byref(self._res)))
        self._keep_host = (vertices, faces)          # the copies may still be reading them
        self.n += 1
        return self._result(self._res)

    def drain(self):
        res = (_lib.PipelineResult * self.depth)()
        n = C.c_int()
        check(lib.meshtok_pipeline_drain(self._h, res, C.byref(n)))
        return [self._result(res[i]) for i in range(n.value)]

    def close(self):
        if getattr(self, "_h", None):
            lib.meshtok_pipeline_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RaggedHostPipeline:
    """HostPipeline for the ragged layout (a CSR of meshes): host buffers in, flat host codes out, `depth` batches in
    flight.  Every slot holds ONE CUDA graph captured on the stated capacities; a submit copies only the live part of the batch
    (vertices, faces, 2 * (max_meshes + 1) offsets), launches the graph and copies back the live code rows - one C call
    (meshtok_pipeline_submit_ragged), no padding on the host, no pad tokens over the bus.

        pipe = model.host_pipeline_ragged(256, 800, 441, 256 * 500, 256 * 300)
        for b in loader:                                   # data.ragged_collate output: pinned host tensors
            done = pipe.submit(b["vertices"], b["faces"], b["face_offsets"], b["vertex_offsets"])   # -> (ticket, codes (sum nf * nvf, Q)) or None
        for ticket, codes in pipe.drain(): ...

    The codes handed back are a view of the slot's pinned buffer (overwritten `depth` submits later)."""

    def __init__(self, model: MeshAutoencoder, max_meshes: int, max_faces: int, max_vertices: int, max_total_faces: int,
                 max_total_vertices: int, depth: int = 2, histogram: Optional[Tensor] = None):
        prep = model._prepare()
        dev = prep["device"]
        model.eval()
        self.model, self.depth, self.nvf, self.Q = model, int(depth), model.num_vertices_per_face, model.num_quantizers
        self.B, self.F, self.V = int(max_meshes), max(int(max_faces), 1), max(int(max_vertices), self.nvf)
        self.cap_faces = min(int(max_total_faces), self.B * self.F)
        self.cap_vertices = min(int(max_total_vertices), self.B * self.V)
        assert self.depth >= 1 and self.B > 0 and self.cap_faces > 0 and self.cap_vertices >= self.nvf
        cap = max(1024, model.edge_slots_per_face * self.cap_faces)
        nbytes = C.c_size_t()
        check(lib.meshtok_tokenize_workspace_bytes_ragged(C.byref(prep["model"]), self.B, self.F, self.V, 0, cap, self.cap_faces, self.cap_vertices,
                                                          C.byref(nbytes)))
        slots = (_lib.PipelineRaggedSlot * self.depth)()
        self._keep, self.outs = [], []
        B1 = self.B + 1
        with torch.cuda.device(dev):
            for s in range(self.depth):
                v = torch.zeros((self.cap_vertices, 3), dtype=torch.float32, device=dev)
                f = torch.zeros((self.cap_faces, self.nvf), dtype=torch.int64, device=dev)
                off = torch.zeros(2 * B1, dtype=torch.int32, device=dev)
                # the batch the graph is captured on: one face in mesh 0, every other mesh empty
                v[: self.nvf] = torch.eye(4, 3, device=dev)[: self.nvf] * 0.5
                f[0] = torch.arange(self.nvf, device=dev)
                off[1:B1] = 1
                off[B1 + 1:] = self.nvf
                codes = torch.empty((self.cap_faces * self.nvf, self.Q), dtype=torch.int64, device=dev)
                ws = torch.zeros(nbytes.value, dtype=torch.uint8, device=dev)
                out = torch.empty(codes.shape, dtype=torch.int64).pin_memory()
                st = torch.zeros(4, dtype=torch.int32).pin_memory()
                oh = torch.zeros(2 * B1, dtype=torch.int32).pin_memory()
                self._keep += [v, f, off, codes, ws, st, oh]
                self.outs.append(out)
                sl = slots[s]
                sl.vertices, sl.faces, sl.offsets, sl.codes = ptr(v), ptr(f), ptr(off), ptr(codes)
                sl.workspace, sl.workspace_bytes = ptr(ws), ws.numel()
                sl.codes_host, sl.status_host, sl.offsets_host = ptr(out), ptr(st), ptr(oh)
            torch.cuda.synchronize(dev)
            handle = C.c_void_p()
            check(lib.meshtok_pipeline_create_ragged(C.byref(prep["model"]), self.B, self.F, self.V, self.cap_faces, self.cap_vertices, int(model.pad_id),
                                                     cap, ptr(histogram), slots, self.depth, C.byref(handle)))
        self._h = handle
        self._histogram = histogram
        self._res = _lib.PipelineResult()
        self._keep_host = [None] * self.depth
        self.n = 0

    def streams(self):
        return [torch.cuda.ExternalStream(lib.meshtok_pipeline_stream(self._h, s)) for s in range(self.depth)]

    def _result(self, r):
        if r.ticket < 0:
            return None
        if r.status:
            MeshAutoencoder._raise_on_status(int(r.status))
        return int(r.ticket), self.outs[r.slot][: int(r.tokens)]

    def submit(self, vertices: Tensor, faces: Tensor, face_offsets, vertex_offsets):
        fo = torch.as_tensor(face_offsets, dtype=torch.int32).contiguous()
        vo = torch.as_tensor(vertex_offsets, dtype=torch.int32).contiguous()
        assert not vertices.is_cuda and not faces.is_cuda and not fo.is_cuda and not vo.is_cuda, "RaggedHostPipeline takes HOST tensors"
        assert vertices.dtype == torch.float32 and faces.dtype == torch.int64 and vertices.is_contiguous() and faces.is_contiguous()
        n = fo.numel() - 1
        assert vo.numel() == n + 1 and vertices.ndim == 2 and vertices.shape[1] == 3 and faces.ndim == 2 and faces.shape[1] == self.nvf
        assert int(fo[-1]) == faces.shape[0] and int(vo[-1]) == vertices.shape[0], "the last offsets must equal the row counts"
        check(lib.meshtok_pipeline_submit_ragged(self._h, ptr(vertices), ptr(faces), ptr(fo), ptr(vo), n, C.byref(self._res)))
        self._keep_host[self.n % self.depth] = (vertices, faces)     # the slot's copies may still be reading them
        self.n += 1
        return self._result(self._res)

    def drain(self):
        res = (_lib.PipelineResult * self.depth)()
        n = C.c_int()
        check(lib.meshtok_pipeline_drain(self._h, res, C.byref(n)))
        return [self._result(res[i]) for i in range(n.value)]

    def close(self):
        if getattr(self, "_h", None):
            lib.meshtok_pipeline_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
