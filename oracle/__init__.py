"""Test infrastructure: CPU restatement of the reference's segmentation path (see unet_oracle.py).
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / --impl reference legs may import this package."""
