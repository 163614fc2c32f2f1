"""CPU stand-ins for the device side of picnic_b200/halo.py: the same pack / unpack-add / mark / pack-leavers / append
contract as the C ABI, on numpy arrays, so that the index-box arithmetic and the message pattern can be exercised with
gloo (or a mailbox) here.  They live with the box-owning CPU workers of the bench's reference arm."""
from oracle.cpu_boxes import NumpyGridBackend, NumpySpeciesBackend  # noqa: F401
