"""Replication strategies for local checkpoints (API mirror of reference ``local/replication/strategies.py``).

``CliqueReplicationStrategy.replicate`` keeps the reference contract (``:88-140``): it returns the state dicts of
*all* clique members (own included, in group-rank order) with host tensors, plus their ids, and leaves the
input hollow.  The payload moves as one packed collective instead of F x N broadcasts (see
``group_utils.GroupWrapper.all_gather_batch``).
"""

import logging
import random
from abc import ABC, abstractmethod
from typing import Generic, List, Mapping, Optional, Sequence, Tuple, TypeVar

import torch

from ...utils import debug_msg, debug_time
from ..base_state_dict import TensorAwareStateDict
from .group_utils import ExchangePlan, GroupWrapper, ProcessGroupLike, parse_group_sequence
from .utils import zip_strict

logger = logging.getLogger(__name__)


class NoReplicasAvailableError(Exception):
    """No rank of the clique holds the requested checkpoint id."""


class ReplicationStrategy(ABC):
    """What a checkpoint manager needs from a replication scheme."""

    @abstractmethod
    def replicate(self, local_ckpt: TensorAwareStateDict, id_: str) -> Tuple[List[TensorAwareStateDict], List[str]]:
        """Exchange ``local_ckpt`` (identified by ``id_``) with the peers; returns the state dicts this rank must
        store and their ids."""

    @abstractmethod
    def retrieve_plan(self, globally_available_ids: Mapping[int, List[str]], wanted: Sequence[str]) -> ExchangePlan:
        """Decide who sends which id to whom, given what every rank holds and what this rank wants."""

    @abstractmethod
    def retrieve_execute(self, *args, **kwargs):
        """Run a plan produced by ``retrieve_plan``."""


class CliqueReplicationStrategy(ReplicationStrategy):
    """Full replication inside one group: every member ends up with every member's shard."""

    def __init__(self, local_group: ProcessGroupLike, target_device="cpu"):
        self.local_group: GroupWrapper = GroupWrapper.wrap(local_group)
        self.target_device = target_device
        self._snapshots: list = []

    def pop_snapshots(self) -> list:
        """Engine snapshot handles created by the last ``replicate`` (the caller waits for / releases them)."""
        snaps, self._snapshots = self._snapshots, []
        return snaps

    @debug_time("CliqueReplicationStrategy.replicate", logger)
    def replicate(self, local_ckpt: TensorAwareStateDict, id_: str) -> Tuple[List[TensorAwareStateDict], List[str]]:
        payload = local_ckpt.pop_tensors()  # local_ckpt is hollow (and picklable) from here on
        with debug_time("all_gather_hollow_ckpt"):
            # skeleton and id travel together (the reference gathers them in two rounds, ``:113`` and ``:135``)
            gathered = self.local_group.all_gather_object((local_ckpt, id_))
        skeletons = [g[0] for g in gathered]
        ids = [g[1] for g in gathered]
        assert all(s.is_hollow for s in skeletons)

        with debug_time("all_gather_others_tensor_data"):
            payloads = self.local_group.all_gather_batch(payload, target_device=self.target_device)
        self._snapshots.extend(self.local_group.last_snapshots)

        sent_bytes = sum(t.nbytes for t in payload)
        recv_bytes = sum(sum(t.nbytes for t in tensors) for tensors in payloads) - sent_bytes
        for skeleton, tensors in zip_strict(skeletons, payloads):
            skeleton.insert_tensors(tensors)
        assert all(not s.is_hollow for s in skeletons)

        debug_msg(f"{sent_bytes=}")
        debug_msg(f"{recv_bytes=}")
        assert local_ckpt.is_hollow
        return skeletons, ids

    @debug_time("CliqueReplicationStrategy.retrieve_plan", logger)
    def retrieve_plan(self, globally_available_ids: Mapping[int, List[str]], wanted: Sequence[str]) -> ExchangePlan:
        """Every rank computes the same plan: a receiver that holds the id serves itself, otherwise a holder is
        drawn with ``random.Random(0)`` from the sorted holders (reference ``:143-179``).

        Raises ``NoReplicasAvailableError`` when nobody in the clique holds a wanted id."""
        rng = random.Random(0)
        with debug_time("all_gather_wanted_ids"):
            wanted_by_rank = self.local_group.all_gather_object(wanted)
        members = self.local_group.ranks
        plan = ExchangePlan(group=self.local_group)
        for receiver, ids in zip(members, wanted_by_rank):
            for wanted_id in ids:
                holders = {r for r in members if wanted_id in globally_available_ids[r]}
                if not holders:
                    raise NoReplicasAvailableError(f"No replicated copies for id={wanted_id} found!")
                sender = receiver if receiver in holders else rng.choice(sorted(holders))
                plan.plan(sender=sender, receiver=receiver, id_=wanted_id)
        return plan

    @debug_time("CliqueReplicationStrategy.retrieve_execute", logger)
    def retrieve_execute(self, *args, **kwargs):
        return self.local_group.execute_plan(*args, **kwargs)

    @classmethod
    @debug_time("CliqueReplicationStrategy.from_replication_params", logger)
    def from_replication_params(
        cls, replication_jump: int = torch.cuda.device_count(), replication_factor: int = 2
    ) -> "CliqueReplicationStrategy":
        """Build the cliques ``n, n+J, ..., n+(F-1)J`` (``J`` = ``replication_jump``, the failure blast radius, e.g.
        GPUs per node; ``F`` = ``replication_factor``) and return the strategy for this rank's clique.

        World size must be a multiple of ``J*F``; e.g. W=32, J=8, F=2 gives 0-8, 1-9, ..., 7-15, 16-24, ..., 23-31.
        Creating the process groups is collective over the world."""
        logger.debug(f"Initializing {cls.__name__}")
        cliques = parse_group_sequence(
            replication_jump=replication_jump,
            replication_factor=replication_factor,
            world_size=torch.distributed.get_world_size(),
        )
        groups = [torch.distributed.new_group(list(ranks)) for ranks in cliques]
        return cls(GroupWrapper.from_list_of_groups(groups), target_device="cpu")


EagerT = TypeVar("EagerT")


class LazyReplicationStrategyBuilder(ReplicationStrategy, ABC, Generic[EagerT]):
    """Defers building the real strategy (which needs process groups) until it is first used."""

    def __init__(self):
        self._replication_strategy: Optional[EagerT] = None

    @property
    def replication_strategy(self) -> EagerT:
        if self._replication_strategy is None:
            self._replication_strategy = self._eager_build()
        return self._replication_strategy

    def replicate(self, local_ckpt: TensorAwareStateDict, id_: str) -> Tuple[List[TensorAwareStateDict], List[str]]:
        return self.replication_strategy.replicate(local_ckpt, id_)

    def retrieve_plan(self, globally_available_ids: Mapping[int, List[str]], wanted: Sequence[str]) -> ExchangePlan:
        return self.replication_strategy.retrieve_plan(globally_available_ids, wanted)

    def retrieve_execute(self, *args, **kwargs):
        return self.replication_strategy.retrieve_execute(*args, **kwargs)

    def pop_snapshots(self) -> list:
        return self.replication_strategy.pop_snapshots()

    @abstractmethod
    def _eager_build(self) -> EagerT:
        """Create the eager strategy."""


class LazyCliqueReplicationStrategy(LazyReplicationStrategyBuilder[CliqueReplicationStrategy]):
    """``CliqueReplicationStrategy.from_replication_params`` evaluated on first use (same parameters)."""

    def __init__(self, replication_jump: int = torch.cuda.device_count(), replication_factor: int = 2):
        super().__init__()
        self.replication_jump = replication_jump
        self.replication_factor = replication_factor

    def _eager_build(self):
        return CliqueReplicationStrategy.from_replication_params(self.replication_jump, self.replication_factor)
