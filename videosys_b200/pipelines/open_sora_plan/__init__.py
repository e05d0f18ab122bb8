from .pipeline_open_sora_plan import (  # noqa: F401
    OpenSoraPlanConfig,
    OpenSoraPlanPipeline,
    OpenSoraPlanV110PABConfig,
    OpenSoraPlanV120PABConfig,
)
