"""Frame partitioning and the map / reduce tasks of the chunked runner.

Mirror of /root/reference/contact_map/frequency_task.py (``block_slices`` :27-47,
``default_slices`` :49-67, ``load_trajectory_task`` :70-88, ``map_task`` :91-106,
``contacts_per_frame_task`` :109-122, ``reduce_all_results`` :125-141).
"""
from .contact_map import ContactFrequency
from .contact_trajectory import _build_contacts
from . import trajectory as _trajectory


def block_slices(n_total, n_per_block):
    """Slices of at most ``n_per_block`` items covering ``range(n_total)``; the remainder,
    if any, becomes one extra (shorter) block."""
    bounds = list(range(0, n_total - n_total % n_per_block, n_per_block))
    slices = [slice(start, start + n_per_block) for start in bounds]
    if n_total % n_per_block:
        slices.append(slice(len(bounds) * n_per_block, n_total))
    return slices


def default_slices(n_total, n_workers):
    """About one block per worker: blocks of ``max(1, n_total // n_workers)`` frames."""
    return block_slices(n_total, max(1, n_total // n_workers))


def load_trajectory_task(subslice, file_name, **kwargs):
    """Load ``file_name`` and return the frames of ``subslice``."""
    return _trajectory.load(file_name, **kwargs)[subslice]


def map_task(subtrajectory, parameters):
    """ContactFrequency of one trajectory segment."""
    return ContactFrequency(subtrajectory, **parameters)


def contacts_per_frame_task(trajectory, contact_object):
    """Per-frame contacts of one segment for a pre-initialised contact object."""
    return _build_contacts(contact_object, trajectory)


def reduce_all_results(contacts):
    """Left fold of ``add_contact_frequency`` over the partial results."""
    accumulator = contacts[0]
    for contact in contacts[1:]:
        accumulator.add_contact_frequency(contact)
    return accumulator
