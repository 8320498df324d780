"""Drop-in package name: ``import qmps.tools`` etc. resolve to the B200 implementation in ``qmps_b200``."""
