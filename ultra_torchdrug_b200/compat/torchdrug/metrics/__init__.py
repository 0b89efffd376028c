"""`torchdrug.metrics` placeholder: imported by reference ultra/task.py:10, not used on the hot path."""
