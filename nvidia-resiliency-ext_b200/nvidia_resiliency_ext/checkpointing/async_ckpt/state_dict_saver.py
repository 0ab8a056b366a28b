"""Planning and finalization halves of an asynchronous ``torch.distributed.checkpoint`` save.

Mirrors the entry points of reference ``checkpointing/async_ckpt/state_dict_saver.py`` (``save_state_dict_async_plan`` ``:236``,
``save_state_dict_async_finalize`` ``:417``, ``CheckpointMetadataCache`` ``:52``, ``init_checkpoint_metadata_cache`` ``:212``, ``verify_global_md_reuse`` ``:374``):
planning (collective; three ways, cheapest first: reuse the cached plan / decentralized planning with a gather only for the
metadata, or none at all when the plans stored in the loaded checkpoint still match / full reduce_scatter) -> ``storage_writer.prepare_write_data`` (stages the tensors; here: one pack kernel + drain) -> the
caller schedules the write -> ``save_state_dict_async_finalize`` gathers write results and lets the coordinator write
``.metadata``.  The optional cache skips the plan exchange when every rank's local plan is unchanged since the last save.
"""

from __future__ import annotations

import logging
from dataclasses import fields
from time import time
from typing import List, Optional, Tuple, Union

import torch
import torch.distributed as dist
from torch.distributed.checkpoint import CheckpointException
from torch.distributed.checkpoint.default_planner import DefaultSavePlanner
from torch.distributed.checkpoint.metadata import STATE_DICT_TYPE, Metadata
from torch.distributed.checkpoint.planner import SavePlan, SavePlanner
from torch.distributed.checkpoint.utils import _DistWrapper, _get_failure_dict

from .core import _collective_device

logger = logging.getLogger(__name__)


class CheckpointMetadataCache:
    """What can be reused between consecutive saves of the same structure: the final (central) plan of this rank, its local
    plan, and -- on the coordinator -- the global metadata of the last save."""

    def __init__(self):
        self.cached_central_plan: Optional[SavePlan] = None
        self.cached_local_plan: Optional[SavePlan] = None
        self.cached_global_metadata: Optional[Metadata] = None
        self.validated_cache_reuse: bool = False
        self.validated_loaded_metadata_reuse: bool = False  # the plans stored in the loaded checkpoint still match ours
        self.loaded_all_plans: Optional[List[SavePlan]] = None

    def set_cached_global_metadata(self, cached_global_metadata: Optional[Metadata]):
        """Seed the cache with metadata loaded from a checkpoint (resume): its per-rank plans allow skipping the first gather."""
        self.cached_global_metadata = cached_global_metadata
        self.loaded_all_plans = getattr(cached_global_metadata, "all_local_plans", None) if cached_global_metadata else None

    def get_cache_metadata(self):
        return self.cached_central_plan, self.cached_local_plan, self.validated_cache_reuse, self.loaded_all_plans

    def set_cache_metadata(self, central_plan: SavePlan, local_plan: SavePlan, global_md_verify_reuse: bool, *, unchanged: Optional[bool] = None):
        """Store the plans of this save.  ``global_md_verify_reuse``: the plans stored in the loaded checkpoint still match
        (reference state_dict_saver.py:115-133).  ``unchanged`` (added, optional): the all-ranks verdict that the local plans
        equal the cached ones; without it the central plans are compared like the reference does."""
        if unchanged is None:
            unchanged = bool(central_plan == self.cached_central_plan)
        self.validated_loaded_metadata_reuse = global_md_verify_reuse
        self.validated_cache_reuse = unchanged and self.cached_central_plan is not None
        self.cached_central_plan = central_plan
        self.cached_local_plan = local_plan

    def prepare_save_state_dict_ret(self, rank: int, coordinator: int, save_state_dict_ret: Tuple) -> Tuple:
        """On the coordinator keep the freshest global metadata and substitute the cached one when planning was skipped."""
        writer, metadata, dist_wrapper = save_state_dict_ret
        if rank == coordinator:
            if metadata is None:  # planning was skipped (cached plan) or needed no metadata exchange (loaded plans match)
                metadata = self.cached_global_metadata
            else:
                self.cached_global_metadata = metadata
        return writer, metadata, dist_wrapper

    def get_metadata_caching_status(self):
        return {"validated_cache_reuse": self.validated_cache_reuse,
                "validated_loaded_metadata_reuse": self.validated_loaded_metadata_reuse,
                "has_central_plan": self.cached_central_plan is not None,
                "has_global_metadata": self.cached_global_metadata is not None}


_checkpoint_metadata_cache: Optional[CheckpointMetadataCache] = None


def init_checkpoint_metadata_cache(cached_global_metadata: Metadata = None):
    """Create (once) the process-wide metadata cache; optionally seed it with metadata loaded from a checkpoint."""
    global _checkpoint_metadata_cache
    if _checkpoint_metadata_cache is None:
        _checkpoint_metadata_cache = CheckpointMetadataCache()
    if cached_global_metadata is not None:
        _checkpoint_metadata_cache.set_cached_global_metadata(cached_global_metadata)
    return _checkpoint_metadata_cache


def get_metadata_caching_status():
    return _checkpoint_metadata_cache.get_metadata_caching_status() if _checkpoint_metadata_cache else None


def _plans_equal(a: Optional[SavePlan], b: Optional[SavePlan]) -> bool:
    if a is None or b is None:
        return False
    return all(getattr(a, f.name) == getattr(b, f.name) for f in fields(a) if f.name != "storage_data")


def _compare_dataclasses(a, b) -> List[str]:
    """Names of the fields in which two plans differ (debug output of the reuse checks)."""
    return [f.name for f in fields(a) if getattr(a, f.name, None) != getattr(b, f.name, None)]


def verify_global_md_reuse(loaded_all_plans: Optional[List[SavePlan]], local_plan: SavePlan, rank: int, dist_wrapper: _DistWrapper) -> bool:
    """Can the global metadata loaded with the checkpoint we resumed from be written again as it is?  Yes when that
    checkpoint stored one local plan per rank of this job and every rank's new local plan equals its stored one (storage
    prefixes aside).  Collective when plans were loaded (one small all_reduce)."""
    if not loaded_all_plans or len(loaded_all_plans) != dist_wrapper.get_world_size():
        return False
    mine = _plans_equal(local_plan, loaded_all_plans[rank])
    if not mine:
        logger.debug(f"rank {rank}: local plan differs from the loaded one in {_compare_dataclasses(local_plan, loaded_all_plans[rank])}")
    return _all_ranks_agree(mine, dist_wrapper)


def _all_ranks_agree(flag: bool, dist_wrapper: _DistWrapper) -> bool:
    if not dist_wrapper.use_dist:
        return flag
    t = torch.tensor([int(flag)], dtype=torch.int, device=_collective_device())
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=dist_wrapper.group)
    return bool(t.item())


def save_state_dict_async_plan(
    state_dict: STATE_DICT_TYPE,
    storage_writer: "FileSystemWriterAsync",  # noqa: F821
    process_group: Optional[dist.ProcessGroup] = None,
    coordinator_rank: int = 0,
    planner: Optional[Union[SavePlanner, DefaultSavePlanner]] = None,
    enable_cache: bool = False,
    metadata_cache: Optional[CheckpointMetadataCache] = None,
) -> Tuple["FileSystemWriterAsync", Union[Metadata, None], _DistWrapper]:  # noqa: F821
    """First stage of an async DCP save: planning + staging.  Collective.  Pass the returned tuple to
    :func:`save_state_dict_async_finalize` once the scheduled write has completed on all ranks."""
    cache = metadata_cache if metadata_cache is not None else (_checkpoint_metadata_cache if enable_cache else None)
    if enable_cache and cache is None:
        cache = init_checkpoint_metadata_cache()
    rank = dist.get_rank() if dist.is_initialized() else 0
    dist_wrapper = _DistWrapper(process_group, dist.is_initialized(), coordinator_rank)
    planner = planner if planner is not None else DefaultSavePlanner()
    global_metadata = None
    t0 = time()

    planner.set_up_planner(state_dict, is_coordinator=dist_wrapper.is_coordinator)
    storage_writer.set_up_storage_writer(dist_wrapper.is_coordinator)
    local_plan = storage_writer.prepare_local_plan(planner.create_local_plan())

    unchanged = loaded_reuse = False
    if cache is not None:
        unchanged = _all_ranks_agree(_plans_equal(local_plan, cache.cached_local_plan), dist_wrapper)
    if unchanged and cache.cached_central_plan is not None:
        logger.debug(f"rank: {rank}, reusing the cached plan")
        central_plan = cache.cached_central_plan  # global metadata comes from the cache on the coordinator
    elif getattr(planner, "can_run_decentralized_global_plan", False) and getattr(storage_writer, "can_run_decentralized_global_plan", False):
        # every rank finishes its own plan; the plans travel (gather, not scatter) only because the coordinator needs them for
        # the global metadata -- and not even that when the checkpoint we resumed from already holds matching ones
        loaded_reuse = verify_global_md_reuse(cache.loaded_all_plans if cache is not None else None, local_plan, rank, dist_wrapper)
        if not loaded_reuse:
            all_local_plans = dist_wrapper.gather_object(local_plan)
            if dist_wrapper.is_coordinator:
                _, global_metadata = planner.create_global_plan(all_local_plans)
                global_metadata.all_local_plans = all_local_plans  # stored in .metadata: lets a resumed job skip this gather
        central_plan = storage_writer.prepare_decentralized_global_plan(planner.create_decentralized_global_plan(local_plan))
    else:
        def global_step(all_local_plans):
            nonlocal global_metadata
            all_local_plans, global_metadata = planner.create_global_plan(all_local_plans)
            return storage_writer.prepare_global_plan(all_local_plans)

        central_plan = dist_wrapper.reduce_scatter("plan", lambda: local_plan, global_step)
    central_plan = planner.finish_plan(central_plan)
    logger.debug(f"rank: {rank}, plan time: {time() - t0}")

    t1 = time()
    storage_writer.prepare_write_data(central_plan, planner)
    logger.debug(f"rank: {rank}, write(async) time: {time() - t1}")
    ret = (storage_writer, global_metadata, dist_wrapper)
    if cache is not None:
        cache.set_cache_metadata(central_plan, local_plan, loaded_reuse, unchanged=unchanged)
        ret = cache.prepare_save_state_dict_ret(rank, coordinator_rank, ret)
    return ret


def save_state_dict_async_finalize(
    storage_writer: "FileSystemWriterAsync", global_metadata: Metadata, dist_wrapper: _DistWrapper  # noqa: F821
) -> None:
    """Second stage: gather every rank's write results, write ``.metadata`` on the coordinator, and raise
    ``CheckpointException`` on *all* ranks if any rank failed (details on the coordinator)."""
    write_results = storage_writer.retrieve_write_results()
    all_results = dist_wrapper.gather_object(write_results)
    failures = {}
    if dist_wrapper.is_coordinator:
        failures = _get_failure_dict(all_results)
        if not failures:
            assert global_metadata is not None
            storage_writer.finish(global_metadata, all_results)
    if dist_wrapper.use_dist:
        flag = torch.tensor([int(bool(failures))], dtype=torch.int, device=_collective_device())
        dist.broadcast(flag, src=dist_wrapper.coordinator_rank, group=dist_wrapper.group)
        failed = bool(flag.item())
    else:
        failed = bool(failures)
    if failed:
        raise CheckpointException("write", failures)
