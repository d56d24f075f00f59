"""ORACLE — test infrastructure only.

A CPU-capable, plain-PyTorch restatement of the reference hot path (AnyV2V i2vgen-xl: I2VGen-XL UNet forward,
DDIM / inverse-DDIM steps, the three PnP hooks, the two sampling loops).  PARITY UNPINNED at the diffusers boundary
(the reference has no tests and its arithmetic lives in the absent pip dependency diffusers==0.26.3); what the
reference does pin — its own hook code and its vendored scheduler — is checked by oracle/make_golden.py against
/root/reference and frozen into tests/golden/.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package.
The product (anyv2v_b200/) never does.
"""
