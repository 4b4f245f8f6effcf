from .dist import (  # noqa: F401
    get_rank, get_world_size, is_main_process, barrier, format_step, mkdir, mkdir_by_main_process,
    init_distributed)
