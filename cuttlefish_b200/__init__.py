"""cuttlefish_b200 -- B200-native block texture encoding behind Cuttlefish's Texture::convert().

Host-side mirror (Python) of the reference's convert surface over the C-ABI in include/cfx.h:

    Texture.convert(format, type, quality, alphaType, colorMask)   lib/include/cuttlefish/Texture.h:740-742
    Texture.data(mip, depth) / dataSize(...)                        lib/include/cuttlefish/Texture.h:780-807

plus array-level helpers `encode()` (host buffers) and `encode_device()` (torch CUDA tensors).
All encoding happens in hand-written sm_100a kernels inside lib/libcfx.so; there is no CPU path.
"""
from .api import (ALPHA, FILTERS, FORMATS, QUALITY, SRC_FORMATS, TYPES, CfxError, ColorMask, Texture,  # noqa: F401
                  block_info, container_header, encode_mip_chain_to_file, encode, encode_batch, encode_device, encode_mip_chain, encode_mip_chain_device, mip_levels, resize, encoded_size, format_is_exact, format_supported, init,
                  init_devices, set_devices, ipc_export, ipc_open, ipc_close, device_count, shutdown, kernel_launches, last_error, shard_block_rows, version)
