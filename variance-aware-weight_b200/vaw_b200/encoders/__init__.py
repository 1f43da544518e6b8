"""Frozen REPA teacher encoders on the B200 library."""
