"""CPU oracle for the EWA-Jinc hot path: TEST INFRASTRUCTURE ONLY (see oracle/jinc_oracle.h).
Nothing under avisynth-jincresize_b200/ may import this package."""
