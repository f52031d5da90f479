"""micromix_b200 -- B200 (sm_100a) drop-in for the hot path of lwy2020/MicroMix.

Only what the path needs: `csrc/` (CUDA kernels + C ABI), `mixedgemm` (the reference's op module),
`qLinearLayer` (the reference's layer), `parallel_utils` (NCCL tensor parallel linears).
"""
__version__ = "0.1.0"
