"""Error types of the runtime (subset of ``tf.errors`` the reference relies on).

``example_between_graph.py:99`` notes that the monitored session "handles
AbortedError in case of preempted PS": :class:`AbortedError` and
:class:`UnavailableError` are the two errors the recoverable session retries on.
"""
__all__ = ["OpError", "FailedPreconditionError", "AbortedError", "UnavailableError", "OutOfRangeError",
           "CancelledError", "DeadlineExceededError", "NotFoundError", "InvalidArgumentError", "DataLossError"]


class OpError(Exception):
    def __init__(self, message="", node_def=None, op=None):
        super().__init__(message)
        self.message, self.node_def, self.op = message, node_def, op


class FailedPreconditionError(OpError):
    pass


class AbortedError(OpError):
    pass


class UnavailableError(OpError):
    pass


class OutOfRangeError(OpError):
    pass


class CancelledError(OpError):
    pass


class DeadlineExceededError(OpError):
    pass


class NotFoundError(OpError):
    pass


class InvalidArgumentError(OpError):
    pass


class DataLossError(OpError):
    """Unrecoverable data corruption (a record or checkpoint whose checksum does not match)."""
