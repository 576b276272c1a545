"""Process-wide distributed environment, same contract as the reference's `breastclip.util.GlobalEnv`
(util/global_env.py:15-34): `.get()` returns a namedtuple (world_size, world_rank, local_rank, num_gpus, master,
summary_writer); the loss modules read rank/world and the TensorBoard writer from it."""
import collections
import os

import torch
import torch.distributed as dist


class SummaryWriter:
    def __init__(self):
        self.train = None
        self.valid = None
        self.global_step = 0


DistEnv = collections.namedtuple("DistEnv", ["world_size", "world_rank", "local_rank", "num_gpus", "master", "summary_writer"])


class GlobalEnv:
    _instance = None

    @staticmethod
    def get():
        if GlobalEnv._instance is None:
            GlobalEnv()
        return GlobalEnv._instance

    @staticmethod
    def reset():
        """Rebuild after init_process_group / destroy_process_group (tests)."""
        GlobalEnv._instance = None
        return GlobalEnv.get()

    def __init__(self):
        if GlobalEnv._instance is not None:
            raise Exception("This class is a singleton")
        if dist.is_available() and dist.is_initialized():
            GlobalEnv._instance = DistEnv(dist.get_world_size(), dist.get_rank(), int(os.environ.get("LOCAL_RANK", 0)), 1,
                                          dist.get_rank() == 0, SummaryWriter())
        else:
            GlobalEnv._instance = DistEnv(1, 0, 0, torch.cuda.device_count(), True, SummaryWriter())
