"""uniter_model/data/data.py: database readers and the base dataset.

Namesake of the reference module: importing it yields `lightningdot_b200.data` itself (same object), so every name the
reference's scripts import from here - private helpers included - is the B200 mirror's."""
import sys

import lightningdot_b200.data as _mirror

sys.modules[__name__] = _mirror
