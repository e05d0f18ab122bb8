from .pipeline_vchitect import VchitectConfig, VchitectPABConfig, VchitectXLPipeline  # noqa: F401
