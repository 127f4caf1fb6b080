"""CPU oracle for the proof-of-burn circuits -- TEST INFRASTRUCTURE ONLY (see oracle/fr.h)."""
