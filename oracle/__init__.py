"""Test infrastructure: CPU oracle for the RSSM hot path (see rssm_oracle.py)."""
