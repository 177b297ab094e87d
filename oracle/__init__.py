"""CPU oracle package -- TEST INFRASTRUCTURE ONLY (see lgdsp_oracle.c header). Never imported by the product."""
