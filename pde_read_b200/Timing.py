"""Wall-clock stopwatch with the reference's interface (reference Code/Timing.py:11-26)."""
import time


class Timer_Exception(Exception):
    pass


class Timer:
    def __init__(self):
        self.Start_Time = None

    def Start(self) -> None:
        self.Start_Time = time.perf_counter()

    def Stop(self) -> float:
        if self.Start_Time is None:
            raise Timer_Exception("Timer has not started. Bad")
        return time.perf_counter() - self.Start_Time
