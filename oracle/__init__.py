"""CPU oracle for the oligo-frequency-vector path — TEST INFRASTRUCTURE, never a product path."""
