"""pose2room_b200 -- B200 (sm_100a) kernels for the P2RNet pose-sequence -> 3-D boxes hot path,
behind the reference's own operator / module API."""
__version__ = "0.1.0"
