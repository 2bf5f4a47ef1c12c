"""CPU oracle for the TVC hot path -- test infrastructure only (see tvc_oracle.h)."""
