"""stub (reference arm only): S3 is unreachable offline"""
def resource(*a, **k):
    raise RuntimeError("boto3 stub: no network")
