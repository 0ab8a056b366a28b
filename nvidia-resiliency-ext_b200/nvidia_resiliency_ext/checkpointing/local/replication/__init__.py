"""Replication of local checkpoint shards inside cliques of ranks."""
