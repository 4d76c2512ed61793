"""Mirror of ``NAMDAnalyzer.lib``: the operator boundary (``pylibFuncs``) and the built
``libnda_b200.so`` live here."""
