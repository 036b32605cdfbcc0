"""dvl/indexer/faiss_indexers.py: DenseIndexer, DenseFlatIndexer, DenseHNSWFlatIndexer.

Namesake of the reference module: importing it yields `lightningdot_b200.indexer` itself (same object), so every name the
reference's scripts import from here - private helpers included - is the B200 mirror's."""
import sys

import lightningdot_b200.indexer as _mirror

sys.modules[__name__] = _mirror
