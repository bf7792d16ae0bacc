"""Exceptions of the CPO hot path; names follow pydrex.exceptions (exceptions.py:6-80)."""


class Error(Exception):
    """Base class for exceptions raised by this package."""


class IterationError(Error):
    """Raised when the on-device integrator fails (reference: minerals.py:457-459)."""

    def __init__(self, message):
        super().__init__(message)
        self.message = message
