"""TEST INFRASTRUCTURE: CPU oracle for the elastic-distance hot path (see elastic_oracle.c)."""
