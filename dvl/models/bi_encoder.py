"""dvl/models/bi_encoder.py: BertEncoder, UniterEncoder, BiEncoder, loss, optimiser helpers.

Namesake of the reference module: importing it yields `lightningdot_b200.bi_encoder` itself (same object), so every name the
reference's scripts import from here - private helpers included - is the B200 mirror's."""
import sys

import lightningdot_b200.bi_encoder as _mirror

sys.modules[__name__] = _mirror
