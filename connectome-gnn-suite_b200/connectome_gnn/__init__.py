"""connectome_gnn - B200-native build of the connectome-gnn-suite message-passing path.

Drop-in for the reference package of the same import name (reference
``connectome_gnn/__init__.py:24-40`` exports the same ten names): put
``connectome-gnn-suite_b200/`` on ``sys.path`` ahead of the reference and existing scripts
(``examples/demo.py``, the README quick start, the reference tests) run unchanged, with collate,
GCN / GraphSAGE layers, BatchNorm/ReLU/dropout, readout, head and loss executing in hand-written
sm_100a kernels behind ``include/cgnn.h``.  No CPU fallback: a CUDA device and the built
``lib/libcgnn.so`` are required for anything that computes.
"""

__version__ = "0.2.0+b200.1"
BACKEND = "cuda-sm_100a"

from .graph import (  # noqa: E402
    ConnectomeBatch,
    ConnectomeDataLoader,
    ConnectomeGraph,
    StreamingStore,
    SubjectStore,
    collate_graphs,
)
from .models import GCNConnectome, GraphSAGEConnectome  # noqa: E402
from .synthetic import REGION_NAMES, generate_connectome, generate_dataset  # noqa: E402
from .train import Trainer  # noqa: E402

__all__ = [
    "ConnectomeGraph", "ConnectomeBatch", "ConnectomeDataLoader", "collate_graphs",
    "generate_connectome", "generate_dataset", "REGION_NAMES",
    "GCNConnectome", "GraphSAGEConnectome", "Trainer",
    "StreamingStore",
    "SubjectStore",
]
