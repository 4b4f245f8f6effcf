from .comm import Comm, SingleComm, TorchComm, FakeComm, make_comm  # noqa: F401
from .ddp import DataParallel, unwrap  # noqa: F401
from .sharded_lamb import ShardedLamb  # noqa: F401
