"""Drop-in mirror of the reference's ``mano`` package (/root/reference/mano/manolayer.py)."""
