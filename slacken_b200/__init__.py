"""slacken_b200 -- B200 (sm_100a) implementation of the Kraken-2-style build/classify hot path of
JNP-Solutions/Slacken, behind the C ABI of include/slacken_gpu.h.

The Python classes here play the role of the Scala host (they mirror the reference's names: Taxonomy, IndexParams,
KeyValueIndex, Classifier, ClassifyParams, KrakenReport) and only move buffers and format text; every computation on
the hot path runs in the CUDA kernels of libslacken_gpu.so."""
from .host import (ClassifiedBatch, Classifier, ClassifyParams, DeviceTimer, GpuContext, IndexParams, KeyValueIndex,
                   LibraryBuilder, ReportCounts, Taxonomy, DEFAULT_TOGGLE_MASK)
from .report import KrakenReport
from ._lib import SlackenGpuError

__all__ = ["ClassifiedBatch", "Classifier", "ClassifyParams", "DeviceTimer", "LibraryBuilder", "GpuContext", "IndexParams", "KeyValueIndex",
           "KrakenReport", "ReportCounts", "Taxonomy", "SlackenGpuError", "DEFAULT_TOGGLE_MASK"]
