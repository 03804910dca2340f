"""oracle/ -- TEST INFRASTRUCTURE: CPU checkers for the CUDA path (never imported by flemingo_b200/)."""
