"""uniter_model/model of the reference: only the KD teacher class name is exposed (see itm.py)."""
