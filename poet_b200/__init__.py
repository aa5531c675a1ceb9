"""poet_b200: B200-native PoET deformable encoder/decoder hot path."""
