"""Inert stand-in for IPython (absent from this image); oracle test infrastructure only."""
