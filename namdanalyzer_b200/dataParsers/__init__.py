"""Mirror of ``NAMDAnalyzer.dataParsers`` for the pair path: ``dcdParser.PairOps``."""
