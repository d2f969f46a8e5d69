"""bpreveal_b200: B200-native (sm_100a) implementation of BPReveal's BPNet training,
prediction and attribution hot path.  See DESIGN.md."""
__version__ = "0.1.0"
