"""``tf.train.Supervisor``: the pre-``MonitoredTrainingSession`` training helper.

The reference's MNIST script is derived from TensorFlow r1.3's ``mnist_replica.py`` (it says so at
``distributed_mnist.py:57``), which -- like most parameter-server programs of that generation -- drives training with a
``Supervisor``: the chief initialises the model or restores it from ``logdir``, the other workers wait until it is
ready, the chief checkpoints in the background, and with ``SyncReplicasOptimizer`` the chief additionally runs
``sync_init_op`` and starts the optimizer's queue runner.  This class provides that protocol on top of the same pieces
``MonitoredTrainingSession`` uses here (``Scaffold`` / ``SessionManager`` / ``Coordinator`` / ``Saver``):

    sv = tf.train.Supervisor(is_chief=is_chief, logdir=train_dir, init_op=init_op, local_init_op=opt.local_step_init_op,
                             ready_for_local_init_op=opt.ready_for_local_init_op, recovery_wait_secs=1, global_step=global_step)
    sess = sv.prepare_or_wait_for_session(server.target, config=config)
    if is_chief and sync:
        sess.run(opt.get_init_tokens_op()); sv.start_queue_runners(sess, [opt.get_chief_queue_runner()])
    while not sv.should_stop(): ... sess.run(train_op, feed_dict=...)
    sv.stop()
"""
from __future__ import annotations

import contextlib
import os
import threading
import time
from typing import Any, Callable, List, Optional, Sequence

from ..framework import errors
from ..framework.variables import get_global_step
from .coordinator import Coordinator, QueueRunner
from .monitored_session import Scaffold, SessionManager

__all__ = ["Supervisor"]


class Supervisor:
    USE_DEFAULT = 0

    def __init__(self, graph=None, ready_op=USE_DEFAULT, ready_for_local_init_op=USE_DEFAULT, is_chief: bool = True,
                 init_op=USE_DEFAULT, init_feed_dict=None, local_init_op=USE_DEFAULT, logdir: Optional[str] = None,
                 summary_op=USE_DEFAULT, saver=USE_DEFAULT, global_step=USE_DEFAULT, save_summaries_secs: float = 120,
                 save_model_secs: float = 600, recovery_wait_secs: float = 30, stop_grace_secs: float = 120,
                 checkpoint_basename: str = "model.ckpt", session_manager=None, summary_writer=USE_DEFAULT, init_fn=None,
                 local_init_run_options=None):
        d = lambda v: None if v is Supervisor.USE_DEFAULT else v          # USE_DEFAULT -> let the Scaffold build it
        self._scaffold = Scaffold(init_op=d(init_op), init_feed_dict=init_feed_dict,
                                  init_fn=(lambda scaffold, sess: init_fn(sess)) if init_fn else None, ready_op=d(ready_op),
                                  ready_for_local_init_op=d(ready_for_local_init_op), local_init_op=d(local_init_op),
                                  saver=d(saver))
        self._is_chief, self._logdir = bool(is_chief), logdir
        self._global_step = get_global_step() if global_step is Supervisor.USE_DEFAULT else global_step
        self._save_model_secs, self._recovery_wait_secs = save_model_secs, recovery_wait_secs
        self._stop_grace_secs = stop_grace_secs
        self._save_path = os.path.join(logdir, checkpoint_basename) if logdir else None
        self._coord = Coordinator()
        self._session_manager = session_manager
        self._threads: List[threading.Thread] = []
        self._queue_runners: List[QueueRunner] = []
        self._sess = None

    # -- properties TF programs read --------------------------------------------------------------------------
    @property
    def is_chief(self) -> bool:
        return self._is_chief

    @property
    def coord(self) -> Coordinator:
        return self._coord

    @property
    def saver(self):
        return self._scaffold.finalize().saver

    @property
    def global_step(self):
        return self._global_step

    @property
    def save_path(self) -> Optional[str]:
        return self._save_path

    @property
    def session_manager(self) -> SessionManager:
        if self._session_manager is None:
            self._session_manager = SessionManager(self._scaffold.finalize(), recovery_wait_secs=self._recovery_wait_secs)
        return self._session_manager

    # -- sessions ---------------------------------------------------------------------------------------------
    def prepare_or_wait_for_session(self, master: str = "", config=None, wait_for_checkpoint: bool = False,
                                    max_wait_secs: float = 7200, start_standard_services: bool = True):
        """Chief: initialise the model or restore the newest checkpoint of ``logdir``; others: poll (every
        ``recovery_wait_secs``) until the chief has done so.  Returns a ready session."""
        self._coord.clear_stop()
        if self._is_chief:
            self._sess = self.session_manager.prepare_session(master, self._logdir, config)
            if start_standard_services:
                self.start_standard_services(self._sess)
        else:
            self._sess = self.session_manager.wait_for_session(master, config, max_wait_secs)
        return self._sess

    @contextlib.contextmanager
    def managed_session(self, master: str = "", config=None, start_standard_services: bool = True,
                        close_summary_writer: bool = True):
        sess = self.prepare_or_wait_for_session(master, config, start_standard_services=start_standard_services)
        try:
            yield sess
        except Exception as e:       # noqa: BLE001 - reported through the coordinator, like TF
            self.request_stop(e)
        finally:
            self.stop(close_summary_writer=close_summary_writer)

    # -- services ---------------------------------------------------------------------------------------------
    def start_standard_services(self, sess) -> List[threading.Thread]:
        """Chief only: a background thread that checkpoints every ``save_model_secs`` (and once more at stop)."""
        if not self._is_chief or not self._save_path or not self._save_model_secs:
            return []
        t = threading.Thread(target=self._checkpoint_loop, args=(sess,), name="dtf-sv-saver", daemon=True)
        self._coord.register_thread(t)
        self._threads.append(t)
        t.start()
        return [t]

    def _save(self, sess) -> None:
        try:
            step = int(sess.run(self._global_step)) if self._global_step is not None else None
            self.saver.save(sess, self._save_path, global_step=step)
        except (errors.OpError, RuntimeError):
            pass                      # the ps / session went away: the restart path handles it

    def _checkpoint_loop(self, sess) -> None:
        while not self._coord.wait_for_stop(self._save_model_secs):
            self._save(sess)
        self._save(sess)

    def start_queue_runners(self, sess, queue_runners: Optional[Sequence[QueueRunner]] = None) -> List[threading.Thread]:
        threads: List[threading.Thread] = []
        for qr in (queue_runners or []):
            threads += qr.create_threads(sess, coord=self._coord, daemon=True, start=True)
            self._queue_runners.append(qr)
        self._threads += threads
        return threads

    def loop(self, timer_interval_secs: float, target: Callable, args: Sequence[Any] = (), kwargs=None) -> threading.Thread:
        def run():
            while not self._coord.wait_for_stop(timer_interval_secs):
                target(*args, **(kwargs or {}))
        t = threading.Thread(target=run, name="dtf-sv-loop", daemon=True)
        self._coord.register_thread(t)
        self._threads.append(t)
        t.start()
        return t

    # -- stopping ---------------------------------------------------------------------------------------------
    def should_stop(self) -> bool:
        return self._coord.should_stop()

    def request_stop(self, ex: Optional[BaseException] = None) -> None:
        self._coord.request_stop(ex)

    def wait_for_stop(self) -> None:
        self._coord.wait_for_stop()

    def stop(self, threads=None, close_summary_writer: bool = True, ignore_live_threads: bool = True) -> None:
        self._coord.request_stop()
        sess = self._sess
        if sess is not None:
            # close the queues this supervisor's runners feed while the session is still usable (their close-on-stop
            # thread races with the session teardown below): consumers on other tasks then drain what is left and end
            # with OutOfRangeError instead of waiting for tokens nobody produces any more
            for qr in self._queue_runners:
                if qr.close_op is not None:
                    try:
                        sess.run(qr.close_op)
                    except Exception:     # noqa: BLE001
                        pass
            try:
                sess.cancel()         # unblock take_grad / token dequeues of this session's queue runners
            except Exception:         # noqa: BLE001
                pass
        try:
            self._coord.join(list(threads or []), stop_grace_period_secs=min(self._stop_grace_secs, 30.0))
        finally:
            if sess is not None:
                sess.close()
                self._sess = None

    # kept for source compatibility with TF programs
    Stop, ShouldStop, RequestStop = stop, should_stop, request_stop
