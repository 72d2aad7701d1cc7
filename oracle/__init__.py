"""CPU oracle of the MH-PPO hot path.  TEST INFRASTRUCTURE -- see oracle/mhppo_oracle.h."""
