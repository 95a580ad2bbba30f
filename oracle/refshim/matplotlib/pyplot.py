"""Import stub (test infrastructure only)."""
