"""
Executes the reference's OWN pure-torch helper functions straight from /root/reference source.

*** TEST INFRASTRUCTURE.  Works only where /root/reference exists (the build container); it is
used by tests/golden/make_golden.py to freeze golden vectors and by the container-only tests that
pin oracle/meshtok_oracle.py.  Nothing is copied: the functions are located with `ast`, their
decorators / annotations are dropped (they need jaxtyping-style `Int['b nf 3']` objects from
meshgpt_pytorch.typing, which cannot be imported without the missing third-party packages) and the
resulting function objects are compiled in memory.

The whole package cannot be imported here: `import meshgpt_pytorch` fails on its first third-party
import (pytorch_custom_utils, einx, torch_geometric, vector_quantize_pytorch, x_transformers ... are
not installed and there is no network).  Only these dependency-free helpers can run as-is:
    data.py:297-340                 derive_face_edges_from_faces
    meshgpt_pytorch.py:94-95        l2norm
    meshgpt_pytorch.py:107-111      pad_at_dim
    meshgpt_pytorch.py:151-153      derive_angle
    meshgpt_pytorch.py:155-183      get_derived_face_features
    meshgpt_pytorch.py:187-201      discretize
    meshgpt_pytorch.py:243-257      scatter_mean
and, for the decode groundwork (oracle/decode_oracle.py):
    meshgpt_pytorch.py:203-218      undiscretize
    meshgpt_pytorch.py:286-379      PixelNorm, SqueezeExcite, Block, ResnetBlock (classes)
"""
from __future__ import annotations

import ast
import os
from types import SimpleNamespace

REFERENCE_ROOT = os.environ.get("MESHTOK_REFERENCE_ROOT", "/root/reference")

_WANTED = {
    "meshgpt_pytorch/data.py": ["derive_face_edges_from_faces"],
    "meshgpt_pytorch/meshgpt_pytorch.py": [
        "exists", "default", "l2norm", "pad_at_dim", "derive_angle",
        "get_derived_face_features", "discretize", "scatter_mean", "undiscretize",
        "PixelNorm", "SqueezeExcite", "Block", "ResnetBlock",
    ],
}


def available() -> bool:
    return all(os.path.isfile(os.path.join(REFERENCE_ROOT, p)) for p in _WANTED)


class _StripTyping(ast.NodeTransformer):
    def visit_FunctionDef(self, node: ast.FunctionDef):
        node.decorator_list = []
        node.returns = None
        for a in node.args.args + node.args.kwonlyargs + node.args.posonlyargs:
            a.annotation = None
        self.generic_visit(node)
        return node


def load() -> SimpleNamespace:
    """Returns a namespace holding the reference's helper functions (compiled from its source)."""
    if not available():
        raise FileNotFoundError(f"{REFERENCE_ROOT} is not present on this machine")
    import torch
    import torch.nn.functional as F
    from einops import rearrange, reduce, repeat
    from math import sqrt
    from einops.layers.torch import Rearrange
    from torch import einsum, nn
    from torch.nn.utils.rnn import pad_sequence

    env = dict(torch=torch, F=F, einsum=einsum, rearrange=rearrange, reduce=reduce, repeat=repeat,
               pad_sequence=pad_sequence, Tensor=torch.Tensor, nn=nn, Module=nn.Module, ModuleList=nn.ModuleList,
               sqrt=sqrt, Rearrange=Rearrange, Tuple=tuple)
    for rel, names in _WANTED.items():
        path = os.path.join(REFERENCE_ROOT, rel)
        with open(path, "r") as fh:
            tree = ast.parse(fh.read(), filename=path)
        picked = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in names]
        missing = set(names) - {n.name for n in picked}
        if missing:
            raise RuntimeError(f"{path}: functions not found: {sorted(missing)}")
        mod = ast.Module(body=[_StripTyping().visit(n) for n in picked], type_ignores=[])
        ast.fix_missing_locations(mod)
        exec(compile(mod, path, "exec"), env)
    return SimpleNamespace(**{n: env[n] for names in _WANTED.values() for n in names})
