"""Memory-pool knob with the interface of adcc/memory_pool.py:31-72 (``adcc.memory_pool``) and
libadcc.AdcMemory (libadcc_src/AdcMemory.cc:43-96, libadcc.pyi:30-66).

On the reference this object picks libtensor's allocator ("standard" or the disk-backed
"libxm"), the block size of the block tensors and the contraction batch size.  The device
backend keeps every tensor in HBM behind torch's caching allocator and runs dense GEMM operands
instead of 16-wide blocks, so the knobs are accepted, validated the way the reference validates
them and recorded, but they do not change the arithmetic.  What this object adds is the HBM
view of the pool (``hbm_stats``)."""
import atexit
import os
import shutil
import tempfile


class AdcMemory:
    def __init__(self):
        self._initialise_called = False
        self._set("", 16, "standard")

    def _set(self, pagefile_directory, max_block_size, allocator):
        if allocator != "standard":
            # AdcMemory.cc:62-66: "libxm" exists only if libtensor was compiled with it
            raise ValueError(f"A memory allocator named '{allocator}' is not known to adcc. "
                             "Perhaps it is not compiled in.")
        self._pagefile_directory = pagefile_directory
        self._max_block_size = int(max_block_size)
        self._allocator = allocator
        self._contraction_batch_size = 21870        # AdcMemory.cc:75

    def initialise(self, pagefile_directory, max_block_size, allocator):
        if self._initialise_called:
            raise ValueError("Cannot initialise AdcMemory object twice.")      # AdcMemory.cc:56-58
        self._set(pagefile_directory, max_block_size, allocator)
        self._initialise_called = True

    @property
    def allocator(self): return self._allocator
    @property
    def max_block_size(self): return self._max_block_size
    @property
    def pagefile_directory(self): return self._pagefile_directory

    @property
    def contraction_batch_size(self): return self._contraction_batch_size

    @contraction_batch_size.setter
    def contraction_batch_size(self, value):
        self._contraction_batch_size = int(value)

    def hbm_stats(self):
        """Bytes of HBM held by the tensor store on the current device."""
        import torch
        if not torch.cuda.is_available():
            return {"allocated": 0, "reserved": 0, "peak_allocated": 0}
        return {"allocated": torch.cuda.memory_allocated(), "reserved": torch.cuda.memory_reserved(),
                "peak_allocated": torch.cuda.max_memory_allocated()}


class MemoryPool(AdcMemory):
    def initialise(self, scratch_directory="/tmp", max_block_size=16, allocator="standard"):
        """adcc/memory_pool.py:32-52."""
        if allocator != "standard":
            raise ValueError(f"A memory allocator named '{allocator}' is not known to adcc. "
                             "Perhaps it is not compiled in.")
        pagefile_directory = tempfile.mkdtemp(prefix="adcc_", dir=scratch_directory)
        try:
            super().initialise(pagefile_directory, max_block_size, allocator)
        except ValueError:
            shutil.rmtree(pagefile_directory, ignore_errors=True)
            raise
        atexit.register(MemoryPool.cleanup, self)

    def cleanup(self):
        if self.pagefile_directory and os.path.isdir(self.pagefile_directory):
            shutil.rmtree(self.pagefile_directory)

    @property
    def scratch_directory(self):
        return os.path.dirname(self.pagefile_directory)

    @property
    def page_files(self):
        return []            # nothing is paged: tensors live in HBM

    @property
    def total_size_page_files(self):
        return 0


memory_pool = MemoryPool()
