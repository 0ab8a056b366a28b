"""Checkpointing sub-package: ``async_ckpt`` (TorchAsyncCheckpoint / AsyncCallsQueue), ``local``
(LocalCheckpointManager, TensorAwareStateDict, clique replication) and ``b200`` (the snapshot engine)."""
