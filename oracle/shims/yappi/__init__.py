"""Import stub (absent offline).  TEST INFRASTRUCTURE ONLY."""
