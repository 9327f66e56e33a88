"""TEST INFRASTRUCTURE ONLY — CPU oracles for the Blight query path.

`oracle.CPort`      plain-C restatement of the reference algorithm (oracle/blight_oracle.c), travels everywhere.
`oracle.Reference`  the reference itself (+ fixes P1/P2) compiled from /root/reference into oracle/_ref/ by
                    oracle/build_ref.sh; available wherever that .so was built or shipped.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package,
and only as the checker or the reported CPU baseline.  Nothing under blight_b200/ imports it.
"""
from .binding import CPort, Reference, build_cport, build_reference, reference_available  # noqa: F401
