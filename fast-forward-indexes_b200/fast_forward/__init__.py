"""fast_forward — B200-native drop-in for the re-ranking hot path of fast-forward-indexes."""
