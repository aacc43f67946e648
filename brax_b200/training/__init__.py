"""RL training around the fused env step."""
