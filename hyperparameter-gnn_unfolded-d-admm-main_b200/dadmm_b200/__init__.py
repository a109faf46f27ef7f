"""dadmm_b200 -- host-side binding of the B200 (sm_100a) unfolded D-ADMM hot path.

``functional`` wraps the C ABI of ``libdadmm_sm100.so`` (``include/dadmm.h``) as raw ops and
``torch.autograd.Function``s; ``graph`` converts ``graph_list`` into the device CSR; ``dist`` shards a
batch of problems over the GPUs of one box.  The reference-compatible ``nn.Module``s live one level
up (``unfolded_DLASSO.py``, ``gnn_dlasso_models_progressive.py``, ...).
"""
from . import _lib, graph, functional  # noqa: F401
from .graph import BatchGraph  # noqa: F401
