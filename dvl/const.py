"""dvl/const.py of the reference: feature widths of the detector outputs."""
IMG_DIM = 2048
IMG_LABEL_DIM = 1601
BUCKET_SIZE = 8192
