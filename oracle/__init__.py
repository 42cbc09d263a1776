"""
CPU oracle for the TASOC prepare-stage sky-background hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``photometry_b200`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do.

PARITY UNPINNED: the reference (``/root/reference``, pure Python) delegates the
arithmetic of this path to photutils 1.3.0, astropy 5.1.0, statsmodels 0.13.2
and Bottleneck 1.3.5, none of which is importable in the build container, and
the reference's own tests hold no numerical golden vectors for the path beyond
four trivial known answers.  This package restates those
upstream algorithms in numpy + the real scipy; the four known answers are
checked in ``tests/test_oracle_kat.py``.
"""
from .backgrounds_oracle import (  # noqa: F401
	FFIImageLite, fit_background, pixel_manual_exclude, move_median_central,
	reduce_mode, kde_density, sigma_clip_bounds, sextractor_background, Background2DOracle,
	radial_geometry, XYCEN, star_mask, star_radius,
)
from .prepare_oracle import (  # noqa: F401
	time_smooth_backgrounds, sumimage_accumulate, prepare_stack,
	TESS_DEFAULT_BITMASK, PIXEL_NOT_USED_FOR_BACKGROUND, PIXEL_MANUAL_EXCLUDE,
)
from . import shenanigans_oracle  # noqa: F401
from .shenanigans_oracle import (  # noqa: F401
	pixel_background_shenanigans, indicator_stack, nan_affected, shuffled_order, mean_shenanigans,
	flag_shenanigans, background_shenanigans, PIXEL_BACKGROUND_SHENANIGANS,
)
from .cube_oracle import load_cube  # noqa: F401
