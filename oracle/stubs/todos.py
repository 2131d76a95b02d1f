"""Empty stand-in for delldu's private `todos` debugging package (unet.py:22)."""
