"""Mirror of ``NAMDAnalyzer.dataAnalysis`` for the pair path: ``RadialDensity``."""
