"""Drop-in for the matcher of /root/reference/mass/utils/experimentation.py (filled in with K4)."""
