"""Exception classes raised at the filter boundary.

Same names, constructor arguments and messages as the reference's
``cguerrero/hydrodem/exceptions.py:6-73`` so that callers' ``except`` clauses
and the reference's own tests (``tests/test_sliding_window.py:78-133``) keep
working against this package.
"""


class HydroDEMException(Exception):
    """Base of every error the conditioning path raises on purpose."""

    def __init__(self, msg=""):
        super().__init__(msg)
        self._msg = msg

    def __str__(self):
        return self._msg


class WindowSizeHighError(HydroDEMException):
    def __init__(self, window_size, grid_dimensions=""):
        super().__init__(f"Window size: {window_size} cannot be higher than grid dimensions: {grid_dimensions}")


class WindowSizeEvenError(HydroDEMException):
    def __init__(self, window_size):
        super().__init__(f"Window size: {window_size} cannot be an even number")


class CenterCloseBorderError(HydroDEMException):
    def __init__(self, center_window, window_size):
        super().__init__(f"Center of window: {center_window} too close of border. Window size: {window_size}")


class NumpyArrayExpectedError(HydroDEMException):
    def __init__(self, provided):
        super().__init__(f"Expected numpy ndarray type. Provided: {type(provided)}")


class InnerSizeError(HydroDEMException):
    # the reference's class takes one argument and reuses the ndarray message (exceptions.py:66-73)
    def __init__(self, provided, window_size=None):
        super().__init__(f"Expected numpy ndarray type. Provided: {type(provided)}")


class DeviceError(HydroDEMException):
    """The CUDA library is missing, no GPU is visible, or a kernel failed.  There is no CPU fallback."""
