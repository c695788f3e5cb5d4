"""Exception types of the reference (`LPException.java`, `SolutionException.java`)."""


class LPException(Exception):
    """lpsolver.LPException — e.g. "This linear program is infeasible" (LPSolver.java:173)."""


class SolutionException(LPException):
    """lpsolver.SolutionException extends LPException — e.g. "This linear program is unbounded"
    (LPSolver.java:105), "Auxiliary lp is unbounded" (:149), "Can't perform degenerate pivot" (:193)."""


class LpsError(RuntimeError):
    """A failure of the native library itself (CUDA error, bad state); carries the status code."""

    def __init__(self, status, message):
        super().__init__("%s (status %d)" % (message, status))
        self.status = status
