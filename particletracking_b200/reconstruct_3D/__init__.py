"""Drop-in for the hot path of the reference's ``ParticleDetection.reconstruct_3D``:
``match2D.match_frame / match_complex / match_csv_complex``,
``tracking.tracking_global_assignment`` and ``matchND.create_weights`` -- same names,
arguments, return values and error behaviour, computed by the sm_100a kernels in
``libptk.so``."""
from . import calibrate_cameras, match2D, matchND, tracking  # noqa: F401
