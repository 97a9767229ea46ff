"""sterne_b200 -- B200-native astrometric log-likelihood hot path of STERNE.

Public surface (mirrors the reference's for this path):
    Gaussianlikelihood, positions                    (simulate.py:325, positions.py:13)
    get_parameters_from_shares, read_inits, readpmparin, read_parfile, ...
The compute is in libsterne_b200.so (csrc/sterne_b200.cu, sm_100a); there is no CPU path.
"""
from .constants import ROOTS, SENTINEL
from .shares import (get_parameters_from_shares, group_elements_by_same_values,
                     parameter_name_to_pmparin_indice, filter_dictionary_of_parameter_with_index,
                     column_table)
from .readers import (readpmparin, create_list_of_dict_VLBI, read_parfile, create_list_of_dict_timing,
                      read_inits, dms2deg, dms_str2deg, decyear2mjd)
from .ephemeris import (AnalyticEarthProvider, TableProvider, NovasDE421Provider, earth_vectors)
from .jpleph import JPLEphemerisProvider
from .problem import CompiledProblem, DeviceContext, MultiDeviceContext, host_register, fp64_peak
from . import summary
from .likelihood import Gaussianlikelihood, positions, positions_batch, set_default_earth_provider

__all__ = [n for n in dir() if not n.startswith('_')]
__version__ = '0.1.0'
