"""Stand-in for nerfacc==0.5.3 (absent offline).  TEST INFRASTRUCTURE ONLY."""
