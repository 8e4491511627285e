"""Drop-in history buffers (mirror of rltime/history/__init__.py:1-11)."""
from .device_history import (DevicePrioritizedReplayHistoryBuffer,
                             DeviceReplayHistoryBuffer)
from .online import OnlineHistoryBuffer

# same class names as the reference so `@python('rltime_b200.history.X')` reads naturally
ReplayHistoryBuffer = DeviceReplayHistoryBuffer
PrioritizedReplayHistoryBuffer = DevicePrioritizedReplayHistoryBuffer


def get_types():
    """Registry entries for the `history` type group (rltime/history/__init__.py:6-11).
    `online` has no device kernel: it is a host-side structure."""
    return {
        "online": OnlineHistoryBuffer,
        "replay": DeviceReplayHistoryBuffer,
        "prioritized_replay": DevicePrioritizedReplayHistoryBuffer,
    }
