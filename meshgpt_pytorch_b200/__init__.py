"""meshgpt_pytorch_b200 - B200-native (sm_100a) mesh tokenization: the hot path of
lucidrains/meshgpt-pytorch's MeshAutoencoder.tokenize / encode / quantize and
derive_face_edges_from_faces behind the reference's own Python signatures.

Importing the package loads libmeshtok.so (hand-written CUDA kernels behind a C ABI, include/meshtok.h);
it raises ImportError if the library has not been built - there is no CPU or PyTorch fallback.
"""
from ._lib import LIB_PATH, MeshtokError  # noqa: F401  (loads the shared library)
from .autoencoder import GraphedRaggedTokenizer, GraphedTokenizer, HostPipeline, MeshAutoencoder, RaggedHostPipeline, ResidualLFQ, ResidualVQ, SAGEConv  # noqa: F401
from .data import (cache_face_edges, derive_face_edges_from_faces, padded_to_ragged, ragged_collate, ragged_slice, ragged_to_padded_codes,  # noqa: F401
                   write_face_edges_cache)
from .decoder import MeshDecoder  # noqa: F401
from .frontend import MeshTransformerFrontEnd  # noqa: F401

__version__ = "0.1.0"
