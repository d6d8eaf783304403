"""ORACLE SHIM — `sounddevice` without PortAudio: no devices, inert InputStream.  Test infrastructure only."""
from unittest.mock import MagicMock


class _Default:
    device = (-1, -1)


default = _Default()
InputStream = MagicMock()


def query_devices(*a, **k):
    return []


def query_hostapis(*a, **k):
    return [{"name": "Mock Host API"}]
