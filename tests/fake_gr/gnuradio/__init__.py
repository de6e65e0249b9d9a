"""Test stub of the `gnuradio` package (see tests/fake_gr/pmt.py)."""
