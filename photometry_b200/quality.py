"""Bit constants of photometry/quality.py that the hot path consumes."""


class TESSQualityFlags:
	"""photometry/quality.py:104-124"""
	AttitudeTweak = 1
	SafeMode = 2
	CoarsePoint = 4
	EarthPoint = 8
	ZeroCrossing = 16
	Desat = 32
	ApertureCosmic = 64
	ManualExclude = 128
	SensitivityDropout = 256
	ImpulsiveOutlier = 512
	CollateralCosmic = 1024
	EarthMoonPlanetInFOV = 2048
	ScatteredLight = 4096
	DEFAULT_BITMASK = (AttitudeTweak | SafeMode | CoarsePoint | EarthPoint
		| Desat | ApertureCosmic | ManualExclude | ScatteredLight)

	@classmethod
	def filter(cls, quality, flags=None):
		"""True where none of ``flags`` is set (quality.py:40-53)."""
		if flags is None:
			flags = cls.DEFAULT_BITMASK
		return (quality & flags) == 0


class PixelQualityFlags:
	"""photometry/quality.py:157-166"""
	NotUsedForBackground = 1
	ManualExclude = 2
	BackgroundShenanigans = 4
	DEFAULT_BITMASK = ManualExclude

	@classmethod
	def filter(cls, quality, flags=None):
		if flags is None:
			flags = cls.DEFAULT_BITMASK
		return (quality & flags) == 0
