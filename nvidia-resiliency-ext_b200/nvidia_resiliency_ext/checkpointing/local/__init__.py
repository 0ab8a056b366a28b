"""Local (node-local, replicated) checkpointing: TensorAwareStateDict contract, LocalCheckpointManager and
clique replication -- API mirror of the reference ``checkpointing/local`` package."""
