"""Stand-in for kaolin==0.14.0 (absent offline).  TEST INFRASTRUCTURE ONLY."""
