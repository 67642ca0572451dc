"""CPU oracle for shortseq_b200 -- TEST INFRASTRUCTURE ONLY (see ssq_oracle.c)."""
