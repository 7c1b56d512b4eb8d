from .tf_presets import get_tf, tex_from_pts  # noqa: F401
from .cameras import in_circles, get_rand_pos  # noqa: F401
from .ingest import volume_from_raw_u8  # noqa: F401
