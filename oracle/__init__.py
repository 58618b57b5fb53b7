"""CPU oracle — test infrastructure only (see hifigan_oracle.py's header)."""
