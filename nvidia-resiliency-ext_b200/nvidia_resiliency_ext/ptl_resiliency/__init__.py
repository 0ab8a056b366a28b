"""PyTorch-Lightning glue for local checkpointing (only the part that calls the checkpoint hot path)."""
