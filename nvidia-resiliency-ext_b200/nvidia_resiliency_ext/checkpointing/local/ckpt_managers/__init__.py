"""Checkpoint managers for node-local checkpoints."""
