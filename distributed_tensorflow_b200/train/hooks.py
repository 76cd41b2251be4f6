"""``SessionRunHook`` protocol and the stock hooks.

Callbacks, in order: ``begin()`` (graph still mutable) ->
``after_create_session(session, coord)`` -> per step ``before_run(ctx)`` (may
return :class:`SessionRunArgs` whose fetches are merged into the step) ->
``after_run(ctx, values)`` (``values.results`` holds this hook's fetches;
``ctx.request_stop()`` ends the loop) -> ``end(session)``.

``StopAtStepHook(num_steps | last_step)`` is what the reference uses
(``example_between_graph.py:64``) and subclasses
(``distributed_mnist.py:41-54,116``): it keeps TF's private attribute names
``_num_steps``, ``_last_step``, ``_global_step_tensor`` because the subclass
touches them.
"""
from __future__ import annotations

import os
import time
from typing import Any, Dict, List, Optional

import numpy as np

from ..framework.variables import get_global_step

__all__ = ["SessionRunHook", "SessionRunArgs", "SessionRunContext", "SessionRunValues", "StopAtStepHook",
           "CheckpointSaverHook", "StepCounterHook", "LoggingTensorHook", "NanTensorHook", "FinalOpsHook",
           "SummarySaverHook", "SecondOrStepTimer", "GlobalStepWaiterHook", "StalenessHook", "ProfilerHook"]


class SessionRunArgs:
    def __init__(self, fetches=None, feed_dict=None, options=None):
        self.fetches, self.feed_dict, self.options = fetches, feed_dict, options


class SessionRunValues:
    def __init__(self, results=None, options=None, run_metadata=None):
        self.results, self.options, self.run_metadata = results, options, run_metadata


class SessionRunContext:
    def __init__(self, original_args: SessionRunArgs, session):
        self._original_args, self._session = original_args, session
        self._stop_requested = False

    @property
    def original_args(self) -> SessionRunArgs:
        return self._original_args

    @property
    def session(self):
        return self._session

    @property
    def stop_requested(self) -> bool:
        return self._stop_requested

    def request_stop(self) -> None:
        self._stop_requested = True


class SessionRunHook:
    def begin(self) -> None:
        pass

    def after_create_session(self, session, coord) -> None:
        pass

    def before_run(self, run_context: SessionRunContext) -> Optional[SessionRunArgs]:
        return None

    def after_run(self, run_context: SessionRunContext, run_values: SessionRunValues) -> None:
        pass

    def end(self, session) -> None:
        pass


class StopAtStepHook(SessionRunHook):
    def __init__(self, num_steps: Optional[int] = None, last_step: Optional[int] = None):
        if num_steps is None and last_step is None:
            raise ValueError("One of num_steps or last_step must be specified.")
        if num_steps is not None and last_step is not None:
            raise ValueError("Only one of num_steps or last_step can be specified.")
        self._num_steps = num_steps
        self._last_step = last_step
        self._global_step_tensor = None

    def begin(self):
        self._global_step_tensor = get_global_step()
        if self._global_step_tensor is None:
            raise RuntimeError("Global step should be created to use StopAtStepHook.")

    def after_create_session(self, session, coord):
        if self._last_step is None:
            global_step = session.run(self._global_step_tensor)
            self._last_step = int(global_step) + self._num_steps

    def before_run(self, run_context):
        return SessionRunArgs(self._global_step_tensor)

    def after_run(self, run_context, run_values):
        # The fetched value is the global step BEFORE this run's update (TF-1.12 semantics): assume the run
        # incremented it, and confirm with a fresh read so that training stops AT last_step, not one step past it.
        global_step = int(run_values.results) + 1
        if global_step >= self._last_step:
            step = int(run_context.session.run(self._global_step_tensor))
            if step >= self._last_step:
                run_context.request_stop()


class SecondOrStepTimer:
    def __init__(self, every_secs: Optional[float] = None, every_steps: Optional[int] = None):
        if (every_secs is None) == (every_steps is None):
            raise ValueError("Exactly one of every_secs and every_steps should be provided.")
        self._every_secs, self._every_steps = every_secs, every_steps
        self.reset()

    def reset(self):
        self._last_time, self._last_step = None, None

    def should_trigger_for_step(self, step: int) -> bool:
        if self._last_step is None:
            return True
        if self._last_step == step:
            return False
        if self._every_secs is not None and time.time() >= self._last_time + self._every_secs:
            return True
        if self._every_steps is not None and step >= self._last_step + self._every_steps:
            return True
        return False

    def update_last_triggered_step(self, step: int):
        now = time.time()
        if self._last_time is None:
            elapsed_t, elapsed_s = None, None
        else:
            elapsed_t, elapsed_s = now - self._last_time, step - self._last_step
        self._last_time, self._last_step = now, step
        return elapsed_t, elapsed_s

    def last_triggered_step(self):
        return self._last_step


class CheckpointSaverHook(SessionRunHook):
    """Chief-only periodic save: at session creation, every ``save_secs``/``save_steps``, and at ``end``."""

    def __init__(self, checkpoint_dir: str, save_secs: Optional[float] = None, save_steps: Optional[int] = None,
                 saver=None, checkpoint_basename: str = "model.ckpt", scaffold=None, listeners=None):
        self._dir = checkpoint_dir
        self._save_path = os.path.join(checkpoint_dir, checkpoint_basename)
        self._saver, self._scaffold = saver, scaffold
        self._timer = SecondOrStepTimer(every_secs=save_secs, every_steps=save_steps)
        self._listeners = listeners or []
        self._global_step_tensor = None
        self.num_saves = 0

    def _get_saver(self):
        if self._saver is not None:
            return self._saver
        if self._scaffold is not None and self._scaffold.saver is not None:
            return self._scaffold.saver
        from .saver import Saver
        self._saver = Saver()
        return self._saver

    def begin(self):
        self._global_step_tensor = get_global_step()
        if self._global_step_tensor is None:
            raise RuntimeError("Global step should be created to use CheckpointSaverHook.")
        self._get_saver()

    def after_create_session(self, session, coord):
        step = int(session.run(self._global_step_tensor))
        self._save(session, step)
        self._timer.update_last_triggered_step(step)

    def before_run(self, run_context):
        return SessionRunArgs(self._global_step_tensor)

    def after_run(self, run_context, run_values):
        step = int(run_values.results)
        if self._timer.should_trigger_for_step(step):
            self._timer.update_last_triggered_step(step)
            self._save(run_context.session, step)

    def end(self, session):
        step = int(session.run(self._global_step_tensor))
        if step != self._timer.last_triggered_step():
            self._save(session, step)

    def _save(self, session, step: int):
        raw = getattr(session, "raw_session", lambda: session)()
        self._get_saver().save(raw, self._save_path, global_step=step)
        self.num_saves += 1


class StepCounterHook(SessionRunHook):
    """Logs ``global_step/sec`` every N steps (TF's MonitoredTrainingSession installs one)."""

    def __init__(self, every_n_steps: int = 100, every_n_secs=None, output_dir=None, summary_writer=None):
        self._timer = SecondOrStepTimer(every_steps=every_n_steps) if every_n_secs is None \
            else SecondOrStepTimer(every_secs=every_n_secs)
        self._writer = summary_writer
        self._global_step_tensor = None
        self.last_steps_per_sec: Optional[float] = None

    def begin(self):
        self._global_step_tensor = get_global_step()

    def before_run(self, run_context):
        return SessionRunArgs(self._global_step_tensor)

    def after_run(self, run_context, run_values):
        step = int(run_values.results)
        if self._timer.should_trigger_for_step(step):
            elapsed_t, elapsed_s = self._timer.update_last_triggered_step(step)
            if elapsed_t:
                self.last_steps_per_sec = elapsed_s / elapsed_t
                print("INFO:dtf:global_step/sec: %g" % self.last_steps_per_sec)
                if self._writer is not None:
                    self._writer.add_scalar("global_step/sec", self.last_steps_per_sec, step)


class LoggingTensorHook(SessionRunHook):
    def __init__(self, tensors, every_n_iter: Optional[int] = None, every_n_secs=None, formatter=None):
        if isinstance(tensors, dict):
            self._tensors = dict(tensors)
        else:
            self._tensors = {getattr(t, "name", str(t)): t for t in tensors}
        self._timer = SecondOrStepTimer(every_steps=every_n_iter) if every_n_secs is None \
            else SecondOrStepTimer(every_secs=every_n_secs)
        self._formatter = formatter
        self._iter = 0
        self._trigger = False

    def before_run(self, run_context):
        self._trigger = self._timer.should_trigger_for_step(self._iter)
        return SessionRunArgs(self._tensors) if self._trigger else None

    def after_run(self, run_context, run_values):
        if self._trigger:
            self._timer.update_last_triggered_step(self._iter)
            if self._formatter:
                print(self._formatter(run_values.results))
            else:
                print("INFO:dtf:" + ", ".join("%s = %s" % (k, v) for k, v in run_values.results.items()))
        self._iter += 1


class NanLossDuringTrainingError(RuntimeError):
    pass


class NanTensorHook(SessionRunHook):
    def __init__(self, loss_tensor, fail_on_nan_loss: bool = True):
        self._loss, self._fail = loss_tensor, fail_on_nan_loss

    def before_run(self, run_context):
        return SessionRunArgs(self._loss)

    def after_run(self, run_context, run_values):
        if np.isnan(run_values.results).any():
            if self._fail:
                raise NanLossDuringTrainingError("NaN loss during training.")
            run_context.request_stop()


class FinalOpsHook(SessionRunHook):
    def __init__(self, final_ops, final_ops_feed_dict=None):
        self._ops, self._feed = final_ops, final_ops_feed_dict
        self.final_ops_values = None

    def end(self, session):
        self.final_ops_values = session.run(self._ops, feed_dict=self._feed)


class SummarySaverHook(SessionRunHook):
    def __init__(self, save_steps=None, save_secs=None, output_dir=None, summary_writer=None, scalars=None):
        self._timer = SecondOrStepTimer(every_secs=save_secs, every_steps=save_steps)
        self._writer, self._dir = summary_writer, output_dir
        self._scalars = dict(scalars or {})
        self._global_step_tensor = None

    def begin(self):
        self._global_step_tensor = get_global_step()
        if self._writer is None and self._dir:
            from ..utils.summary import FileWriter
            self._writer = FileWriter(self._dir)

    def before_run(self, run_context):
        fetch = {"__step": self._global_step_tensor}
        fetch.update(self._scalars)
        return SessionRunArgs(fetch)

    def after_run(self, run_context, run_values):
        res = run_values.results
        step = int(res["__step"])
        if self._writer is not None and self._timer.should_trigger_for_step(step):
            self._timer.update_last_triggered_step(step)
            for k, v in res.items():
                if k != "__step":
                    self._writer.add_scalar(k, float(v), step)
            self._writer.flush()


class GlobalStepWaiterHook(SessionRunHook):
    """Delay a worker until the global step reaches ``wait_until_step`` (staggered async start)."""

    def __init__(self, wait_until_step: int):
        self._wait, self._done, self._gs = wait_until_step, False, None

    def begin(self):
        self._gs = get_global_step()

    def before_run(self, run_context):
        if self._done or self._wait <= 0:
            self._done = True
            return None
        while int(run_context.session.run(self._gs)) < self._wait:
            time.sleep(0.05)
        self._done = True
        return None


class StalenessHook(SessionRunHook):
    """Measures async-replica staleness per step (BASELINE.json config 4).

    staleness = (global_step right after this worker's apply) - (global_step this worker pulled
    parameters at) - 1, i.e. how many *other* updates landed between pull and apply.  The pull-time
    step is read in ``before_run`` (its own ``run``), the apply-time step is fetched with the train op.
    """

    def __init__(self):
        self._gs = None
        self._pulled = 0
        self.samples: List[int] = []

    def begin(self):
        self._gs = get_global_step()

    def before_run(self, run_context):
        self._pulled = int(run_context.session.run(self._gs))
        return SessionRunArgs(self._gs)

    def after_run(self, run_context, run_values):
        self.samples.append(max(0, int(run_values.results) - self._pulled - 1))

    def histogram(self) -> Dict[int, int]:
        out: Dict[int, int] = {}
        for s in self.samples:
            out[s] = out.get(s, 0) + 1
        return dict(sorted(out.items()))

    def mean(self) -> float:
        return float(np.mean(self.samples)) if self.samples else 0.0


class ProfilerHook(SessionRunHook):
    """``tf.train.ProfilerHook``: every ``save_steps`` global steps (or ``save_secs`` seconds) run ONE step with
    ``RunOptions(trace_level=FULL_TRACE)`` and write its chrome trace to ``<output_dir>/timeline-<step>.json`` (one
    process per ``/job/task`` device) -- the per-run tracing of the reference's in-graph examples
    (``example_in_graph.py:42-43,65-68``) made available inside a training loop."""

    def __init__(self, save_steps: Optional[int] = None, save_secs: Optional[float] = None, output_dir: str = "",
                 show_dataflow: bool = True, show_memory: bool = False):
        self._timer = SecondOrStepTimer(every_secs=save_secs, every_steps=save_steps)
        self._dir, self._show_dataflow, self._show_memory = output_dir, show_dataflow, show_memory
        self._global_step_tensor = None
        self._next_step = None
        self._tracing = False
        self.files: List[str] = []

    def begin(self):
        self._global_step_tensor = get_global_step()
        if self._global_step_tensor is None:
            raise RuntimeError("Global step should be created to use ProfilerHook.")

    def before_run(self, run_context):
        self._tracing = self._next_step is None or self._timer.should_trigger_for_step(self._next_step)
        opts = None
        if self._tracing:
            from ..client.session import RunOptions
            opts = RunOptions(trace_level=RunOptions.FULL_TRACE)
        return SessionRunArgs(self._global_step_tensor, options=opts)

    def after_run(self, run_context, run_values):
        step = int(run_values.results) + 1
        if self._tracing and run_values.run_metadata is not None:
            from ..utils.timeline import Timeline
            self._timer.update_last_triggered_step(step)
            os.makedirs(self._dir or ".", exist_ok=True)
            path = os.path.join(self._dir or ".", "timeline-%d.json" % step)
            with open(path, "w") as f:
                f.write(Timeline(run_values.run_metadata.step_stats).generate_chrome_trace_format(
                    show_dataflow=self._show_dataflow, show_memory=self._show_memory))
            self.files.append(path)
        self._next_step = step + 1
