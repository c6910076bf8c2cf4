"""Logger of the package -- same name and level as the reference (probreg/log.py:1-6) so that
``logging.getLogger("probreg")`` configuration keeps working after the switch."""
import logging

log = logging.getLogger("probreg")
log.setLevel(logging.INFO)
if not log.handlers:
    log.addHandler(logging.StreamHandler())
