"""CPU oracle -- TEST INFRASTRUCTURE ONLY (see oracle/oracle_int2.cpp header)."""
