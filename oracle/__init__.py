"""TEST INFRASTRUCTURE.  CPU oracle for the Sheet03 two-stream path (a restatement of the reference on stock
PyTorch fp32 + numpy).  Import only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs; the product package video_analytics_b200 never imports this."""
