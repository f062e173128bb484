"""dfd_starter_b200 — B200-native finite-difference learner hot path with the
reference's (nexus-rl/dfd-starter) object interface.  Importing the package needs
neither the shared library nor a GPU; using it needs both (no CPU fallback)."""
from .fd_return import FDReturn, ReturnBatch
from .fd_state import FDState
from .dsgd import DSGD
from .noise_sources import SharedNoiseTable, RNGNoiseSource, SimpleNoiseSource
from .finite_differences import FiniteDifferences
from .worker import Worker, SyntheticAgent
from .grpc_worker import GRPCWorker, RPCServer, RPCClient
from . import wire
from .obs_stats import WelfordRunningStat, normalize_obs, member_obs_stats
from .strategy import StrategyHandler, SparseHistoryManager, StrategyPoint, strategy_distances
from .policies import MujocoPolicy, DiscretePolicy, AtariPolicy, ImpalaPolicy, Policy

__all__ = ["FDReturn", "ReturnBatch", "FDState", "DSGD", "SharedNoiseTable", "RNGNoiseSource", "SimpleNoiseSource", "FiniteDifferences", "Worker", "SyntheticAgent", "GRPCWorker", "RPCServer", "RPCClient", "wire", "WelfordRunningStat", "normalize_obs", "member_obs_stats", "StrategyHandler", "SparseHistoryManager", "StrategyPoint", "strategy_distances",
           "MujocoPolicy", "DiscretePolicy", "AtariPolicy", "ImpalaPolicy", "Policy"]
