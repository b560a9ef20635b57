"""hash_join_codes_knl_b200 -- B200-native hash joins (NPJ, PHJ, CPRA) behind the reference's
interface.  The work is done by hand-written sm_100a kernels in libhjb200.so (csrc/); this
package is the thin host-side binding the tests and bench.py drive it through, plus the
torch.distributed plumbing of the multi-GPU CPRA exchange.  No CPU fallback."""
from .api import Engine, JoinResult, HjbError  # noqa: F401
from . import datagen  # noqa: F401
