"""TEST INFRASTRUCTURE ONLY (see oracle/sfm_oracle.c header).  Importable from tests/, smoke() and bench.py's cpu_baseline leg."""
