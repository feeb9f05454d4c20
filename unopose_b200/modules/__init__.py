"""State-dict compatible counterparts of the reference's matching modules:
the layers around the hot path stay PyTorch host code; the hot-path calls go to the sm_100a kernels."""
from .matching import CoarsePointMatchingOneRef, FinePointMatchingOneRef, PositionalEncoding  # noqa: F401
from .transformer import (GeometricStructureEmbedding, GeometricTransformer, LinearTransformerLayer,  # noqa: F401
                          SparseToDenseTransformer)
