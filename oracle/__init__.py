"""CPU oracle (test infrastructure only; never imported by speech_separation_b200)."""
