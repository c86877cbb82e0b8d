"""TEST INFRASTRUCTURE: the CPU checkers (see oracle/refbind.py, oracle/ref_harness.cpp, oracle/pt_oracle.c)."""
