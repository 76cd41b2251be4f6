"""``dtf.train``: cluster runtime, optimizers, sync replicas, monitored sessions, checkpoints."""
from ..framework.device import replica_device_setter
from ..framework.variables import create_global_step, get_global_step, get_or_create_global_step
from ..parallel.cluster import ClusterSpec
from ..parallel.server import Server
from .coordinator import Coordinator, QueueRunner
from .hooks import (CheckpointSaverHook, FinalOpsHook, GlobalStepWaiterHook, LoggingTensorHook, NanTensorHook, ProfilerHook,
                    SecondOrStepTimer, SessionRunArgs, SessionRunContext, SessionRunHook, SessionRunValues,
                    StalenessHook, StepCounterHook, StopAtStepHook, SummarySaverHook)
from .monitored_session import (ChiefSessionCreator, MonitoredSession, MonitoredTrainingSession, Scaffold,
                                SessionManager, SingularMonitoredSession, WorkerSessionCreator)
from .optimizer import (AdagradOptimizer, AdamOptimizer, GradientDescentOptimizer, MomentumOptimizer, Optimizer, RMSPropOptimizer,
                        cosine_decay, exponential_decay, inverse_time_decay, natural_exp_decay, piecewise_constant, polynomial_decay)
from .saver import (CheckpointState, NewCheckpointReader, Saver, checkpoint_exists, get_checkpoint_state,
                    latest_checkpoint, list_variables, load_checkpoint, update_checkpoint_state)
from .sync_replicas import SyncReplicasOptimizer, SyncReplicasOptimizerHook
from .supervisor import Supervisor
from .extras import (AdadeltaOptimizer, ExponentialMovingAverage, add_queue_runner, global_step, init_from_checkpoint, load_variable,
                     start_queue_runners, write_graph)
