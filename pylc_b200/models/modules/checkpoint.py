"""
Import location of the reference's Checkpoint (reference models/modules/checkpoint.py:17-66).  The class lives next
to Model in pylc_b200/models/model.py; `from models.modules.checkpoint import Checkpoint` keeps working after the
package switch described in INTEGRATION.md section A.
"""
from ..model import Checkpoint, strip_module_prefix  # noqa: F401
