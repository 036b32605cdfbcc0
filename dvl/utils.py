"""dvl/utils.py: loss wrapper, rank helpers, retrieve_query, get_model_encoded_vecs.

Namesake of the reference module: importing it yields `lightningdot_b200.utils` itself (same object), so every name the
reference's scripts import from here - private helpers included - is the B200 mirror's."""
import sys

import lightningdot_b200.utils as _mirror

sys.modules[__name__] = _mirror
