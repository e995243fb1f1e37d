"""gguf-b200: a B200-native GGUF decode engine behind the llama-gguf-inference gateway.

The directory is named `llama-gguf-inference_b200` (not an importable identifier); import it as
`ggufb200` (the shim package at the repository root points its __path__ here).

Host side only holds what the decode path needs: GGUF reader, tokenizer, scheduler, HTTP boundary, and the
ctypes binding of the C-ABI library built from csrc/ (`libggufb200.so`).  There is no CPU fallback: every
compute entry point raises if the CUDA library is missing.
"""
__version__ = "0.1.0"
