"""``MonitoredTrainingSession`` and friends.

Behaviour reproduced (reference ``distributed_mnist.py:144-152``,
``example_between_graph.py:92-107``):

* **chief** (``is_chief=True``): restore the latest checkpoint in
  ``checkpoint_dir`` if there is one, otherwise run the init ops; installs a
  :class:`CheckpointSaverHook` (``save_checkpoint_secs``, default 600 s; also
  saves at session creation and at ``end``), a :class:`StepCounterHook`
  (every 100 steps) and, when summaries are requested, a summary saver;
* **non-chief**: polls until the chief has initialised every global variable
  ("non-chief waits for chief");
* ``run()`` = hooks' ``before_run`` -> ONE merged ``Session.run`` -> hooks'
  ``after_run``; ``should_stop()``; leaving the ``with`` block runs ``end()``
  hooks and closes;
* **recoverable**: on ``AbortedError`` / ``UnavailableError`` (a preempted ps,
  reference comment at ``example_between_graph.py:99``) the session is
  rebuilt -- the chief re-restores from the last checkpoint, workers re-wait --
  ``after_create_session`` hooks run again and the step is retried.
"""
from __future__ import annotations

import os
import time
from typing import Any, Dict, List, Optional, Sequence

from ..client.session import ConfigProto, Session
from ..framework import errors
from ..framework import ops as _ops
from ..framework.graph import GraphKeys, get_default_graph
from ..framework.variables import (global_variables, global_variables_initializer, local_variables,
                                   local_variables_initializer, report_uninitialized_variables,
                                   get_global_step)
from .coordinator import Coordinator
from .hooks import (CheckpointSaverHook, SessionRunArgs, SessionRunContext, SessionRunHook, SessionRunValues,
                    StepCounterHook, SummarySaverHook)

__all__ = ["MonitoredTrainingSession", "MonitoredSession", "SingularMonitoredSession", "Scaffold",
           "ChiefSessionCreator", "WorkerSessionCreator", "SessionManager"]

_RECOVERABLE = (errors.AbortedError, errors.UnavailableError)


class Scaffold:
    """Holds the init / ready / saver pieces; builds defaults on ``finalize()``."""

    def __init__(self, init_op=None, init_feed_dict=None, init_fn=None, ready_op=None,
                 ready_for_local_init_op=None, local_init_op=None, summary_op=None, saver=None):
        self.init_op, self.init_feed_dict, self.init_fn = init_op, init_feed_dict, init_fn
        self.ready_op, self.ready_for_local_init_op = ready_op, ready_for_local_init_op
        self.local_init_op, self.summary_op, self.saver = local_init_op, summary_op, saver
        self._finalized = False

    def finalize(self) -> "Scaffold":
        if self._finalized:
            return self
        if self.init_op is None:
            self.init_op = global_variables_initializer()
        if self.ready_op is None:
            self.ready_op = report_uninitialized_variables(global_variables())
        if self.ready_for_local_init_op is None:
            self.ready_for_local_init_op = report_uninitialized_variables(global_variables())
        if self.local_init_op is None:
            self.local_init_op = local_variables_initializer()
        if self.saver is None:
            from .saver import Saver
            self.saver = Saver(allow_empty=True)
        self._finalized = True
        return self


class SessionManager:
    """Create a ready session: chief initialises / restores, workers wait (TF ``SessionManager``)."""

    def __init__(self, scaffold: Scaffold, recovery_wait_secs: float = 0.5):
        self._scaffold, self._wait = scaffold, recovery_wait_secs

    def prepare_session(self, master: str, checkpoint_dir: Optional[str], config=None) -> Session:
        sc = self._scaffold
        sess = Session(master, config=config)
        restored = False
        if checkpoint_dir:
            from .saver import latest_checkpoint
            path = latest_checkpoint(checkpoint_dir)
            if path:
                sc.saver.restore(sess, path)
                restored = True
        if not restored:
            sess.run(sc.init_op, feed_dict=sc.init_feed_dict)
            if sc.init_fn:
                sc.init_fn(sc, sess)
        else:
            missing = sess.run(sc.ready_op)
            if missing:
                # variables created after the checkpoint was written (e.g. new optimizer slots)
                g = get_default_graph()
                sess.run(_ops.group(*[g.variables[n].initializer for n in missing if n in g.variables]))
        sess.run(sc.local_init_op)
        not_ready = sess.run(sc.ready_op)
        if not_ready:
            raise RuntimeError("Init operations did not make model ready. Uninitialised: %s" % (not_ready,))
        return sess

    def wait_for_session(self, master: str, config=None, max_wait_secs: float = 7200.0) -> Session:
        sc = self._scaffold
        deadline = time.time() + max_wait_secs
        sess = Session(master, config=config)
        while True:
            try:
                not_ready = sess.run(sc.ready_for_local_init_op)
            except _RECOVERABLE:
                not_ready = ["<ps unavailable>"]
            if not not_ready:
                sess.run(sc.local_init_op)
                return sess
            if time.time() > deadline:
                sess.close()
                raise errors.DeadlineExceededError("Session was not ready after waiting %d secs." % max_wait_secs)
            time.sleep(self._wait)


class ChiefSessionCreator:
    def __init__(self, scaffold=None, master: str = "", config=None, checkpoint_dir=None,
                 checkpoint_filename_with_path=None):
        self._scaffold = scaffold or Scaffold()
        self._master, self._config, self._dir = master, config, checkpoint_dir

    def create_session(self) -> Session:
        self._scaffold.finalize()
        return SessionManager(self._scaffold).prepare_session(self._master, self._dir, self._config)


class WorkerSessionCreator:
    def __init__(self, scaffold=None, master: str = "", config=None, max_wait_secs: float = 30 * 60):
        self._scaffold = scaffold or Scaffold()
        self._master, self._config, self._max_wait = master, config, max_wait_secs

    def create_session(self) -> Session:
        self._scaffold.finalize()
        return SessionManager(self._scaffold).wait_for_session(self._master, self._config, self._max_wait)


class MonitoredSession:
    def __init__(self, session_creator=None, hooks: Optional[Sequence[SessionRunHook]] = None,
                 should_recover: bool = True, stop_grace_period_secs: float = 120):
        self._hooks: List[SessionRunHook] = list(hooks or [])
        self._creator = session_creator or ChiefSessionCreator()
        self._should_recover = should_recover
        self._grace = stop_grace_period_secs
        self._graph = get_default_graph()
        for h in self._hooks:
            h.begin()
        self._coord = Coordinator()
        self._sess: Optional[Session] = None
        self._stop_requested = False
        self._closed = False
        self.num_recoveries = 0
        self._create()

    # -- session (re)creation ------------------------------------------------------------------------
    def _create(self) -> None:
        while True:
            try:
                self._coord.clear_stop()
                self._sess = self._creator.create_session()
                for h in self._hooks:
                    h.after_create_session(self._sess, self._coord)
                return
            except _RECOVERABLE as e:
                if not self._should_recover:
                    raise
                print("INFO:dtf:An error was raised while a session was being created "
                      "(%s); retrying." % type(e).__name__)
                self._teardown_session()
                time.sleep(0.5)

    def _teardown_session(self) -> None:
        self._coord.request_stop()
        if self._sess is not None:
            try:
                self._sess.cancel()
            except Exception:
                pass
        try:
            self._coord.join(stop_grace_period_secs=min(self._grace, 5.0))
        except Exception:
            pass
        if self._sess is not None:
            try:
                self._sess.close()
            except Exception:
                pass
            self._sess = None

    def raw_session(self) -> Session:
        return self._sess

    _tf_sess = raw_session

    @property
    def graph(self):
        return self._graph

    # -- running -----------------------------------------------------------------------------------------
    def run(self, fetches, feed_dict=None, options=None, run_metadata=None):
        if self._closed:
            raise RuntimeError("Attempted to use a closed MonitoredSession.")
        while True:
            try:
                return self._run_once(fetches, feed_dict, options, run_metadata)
            except _RECOVERABLE as e:
                if not self._should_recover:
                    raise
                print("INFO:dtf:An error was raised (%s). This may be due to a preemption of a parameter "
                      "server; the session is being recovered." % type(e).__name__)
            except errors.FailedPreconditionError as e:
                # a restarted ps lost its variables: same recovery as an aborted session
                if not self._should_recover or "uninitialized" not in str(e):
                    raise
                print("INFO:dtf:parameter server lost state (%s); recovering." % str(e).splitlines()[0])
            self.num_recoveries += 1
            self._teardown_session()
            self._create()

    def _run_once(self, fetches, feed_dict, options, run_metadata):
        ctx = SessionRunContext(SessionRunArgs(fetches, feed_dict), self._sess)
        merged_feed = dict(feed_dict or {})
        hook_fetches: Dict[int, Any] = {}
        for i, h in enumerate(self._hooks):
            req = h.before_run(ctx)
            if req is not None:
                if req.fetches is not None:
                    hook_fetches[i] = req.fetches
                if req.feed_dict:
                    for k in req.feed_dict:
                        if k in merged_feed:
                            raise RuntimeError("Same tensor is fed by two hooks.")
                    merged_feed.update(req.feed_dict)
                if req.options is not None and options is None:
                    options = req.options
        if options is not None and run_metadata is None and getattr(options, "trace_level", 0):
            from ..client.session import RunMetadata
            run_metadata = RunMetadata()          # a hook asked for a trace: give after_run something to read
        out = self._sess.run({"caller": fetches, "hooks": hook_fetches}, feed_dict=merged_feed or None,
                             options=options, run_metadata=run_metadata)
        for i, h in enumerate(self._hooks):
            h.after_run(ctx, SessionRunValues(results=out["hooks"].get(i), options=options,
                                              run_metadata=run_metadata))
        if ctx.stop_requested:
            self._stop_requested = True
        return out["caller"]

    def should_stop(self) -> bool:
        if self._closed or self._stop_requested:
            return True
        if self._coord.should_stop():
            return True
        return False

    def request_stop(self) -> None:
        self._stop_requested = True

    # -- closing -------------------------------------------------------------------------------------------
    def close(self) -> None:
        self._close_internal(None)

    def _close_internal(self, exc_type) -> None:
        if self._closed:
            return
        try:
            if exc_type is None or exc_type in (errors.OutOfRangeError, StopIteration):
                for h in self._hooks:
                    h.end(self._sess)
        finally:
            self._closed = True
            self._teardown_session()

    def __enter__(self) -> "MonitoredSession":
        return self

    def __exit__(self, exc_type, exc, tb) -> bool:
        self._close_internal(exc_type)
        # an OutOfRange (input exhausted) or StopIteration ends training cleanly
        return exc_type in (errors.OutOfRangeError, StopIteration)


class SingularMonitoredSession(MonitoredSession):
    def __init__(self, hooks=None, scaffold=None, master="", config=None, checkpoint_dir=None,
                 stop_grace_period_secs=120):
        super().__init__(ChiefSessionCreator(scaffold, master, config, checkpoint_dir), hooks,
                         should_recover=False, stop_grace_period_secs=stop_grace_period_secs)


def MonitoredTrainingSession(master: str = "", is_chief: bool = True, checkpoint_dir: Optional[str] = None,
                             scaffold: Optional[Scaffold] = None, hooks: Optional[Sequence[SessionRunHook]] = None,
                             chief_only_hooks: Optional[Sequence[SessionRunHook]] = None,
                             save_checkpoint_secs: Optional[float] = 600, save_summaries_steps: Optional[int] = 100,
                             save_summaries_secs=None, config: Optional[ConfigProto] = None,
                             stop_grace_period_secs: float = 120, log_step_count_steps: Optional[int] = 100,
                             max_wait_secs: float = 7200, save_checkpoint_steps: Optional[int] = None,
                             summary_dir: Optional[str] = None) -> MonitoredSession:
    scaffold = scaffold or Scaffold()
    all_hooks: List[SessionRunHook] = list(hooks or [])
    if not is_chief:
        creator = WorkerSessionCreator(scaffold, master, config, max_wait_secs)
        return MonitoredSession(creator, all_hooks, stop_grace_period_secs=stop_grace_period_secs)
    if chief_only_hooks:
        all_hooks.extend(chief_only_hooks)
    creator = ChiefSessionCreator(scaffold, master, config, checkpoint_dir)
    has_gs = get_global_step() is not None
    if checkpoint_dir and has_gs:
        if log_step_count_steps and log_step_count_steps > 0:
            all_hooks.append(StepCounterHook(every_n_steps=log_step_count_steps))
        summaries = get_default_graph().get_collection(GraphKeys.SUMMARIES)
        if summaries and (save_summaries_steps or save_summaries_secs):
            all_hooks.append(SummarySaverHook(save_steps=save_summaries_steps if not save_summaries_secs else None,
                                              save_secs=save_summaries_secs,
                                              output_dir=summary_dir or checkpoint_dir,
                                              scalars={s.tag: s.tensor for s in summaries}))
        if (save_checkpoint_secs and save_checkpoint_secs > 0) or (save_checkpoint_steps and save_checkpoint_steps > 0):
            all_hooks.append(CheckpointSaverHook(
                checkpoint_dir, save_secs=save_checkpoint_secs if not save_checkpoint_steps else None,
                save_steps=save_checkpoint_steps, scaffold=scaffold))
    elif log_step_count_steps and log_step_count_steps > 0 and has_gs and checkpoint_dir is None:
        pass
    return MonitoredSession(creator, all_hooks, stop_grace_period_secs=stop_grace_period_secs)
