"""Host cores a test process may actually use (cgroup quota aware): see oracle/synth_fast.py::usable_cores."""
from oracle.synth_fast import usable_cores  # noqa: F401
