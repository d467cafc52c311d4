"""TEST INFRASTRUCTURE ONLY — see oracle/canny_oracle.c.  Never imported by canny_edge_b200/."""
