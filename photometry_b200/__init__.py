"""
photometry_b200 -- B200-native (sm_100a) implementation of the TASOC prepare-stage sky-background hot
path: ``photometry.backgrounds.fit_background`` over a CCD's FFI stack, background time smoothing and
the sumimage accumulation.  Host side: Python + a C-ABI CUDA library (include/tbk.h).
"""
from .backgrounds import fit_background, BackgroundFitter, make_meta, meta_from_headers  # noqa: F401
from .io import FFIImage  # noqa: F401
from .quality import TESSQualityFlags, PixelQualityFlags  # noqa: F401
from .prepare import prepare_stack, fit_stack_host, SectorResult  # noqa: F401
from .ingest import load_ffi_stack, decode_ffi_be  # noqa: F401

__version__ = '0.1.0'
from .shenanigans import background_shenanigans, shenanigans_indicator, mean_shenanigans, flag_shenanigans  # noqa: F401
from .pixel_flags import pixel_background_shenanigans  # noqa: F401
from .stamps import StampServer  # noqa: F401
