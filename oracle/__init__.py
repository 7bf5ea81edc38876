"""TEST INFRASTRUCTURE ONLY - CPU oracle for hevcasm_b200 (see hevc_oracle.h).  Import from tests/, smoke() and
bench.py's cpu_baseline / --impl reference legs only."""
