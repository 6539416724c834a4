"""CPU oracle for pypairs_b200 -- TEST INFRASTRUCTURE ONLY (see oracle/pp_oracle.c)."""
