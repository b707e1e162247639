from dampr_b200.inputs import *  # noqa: F401,F403
from dampr_b200.inputs import read_paths, PathInput, TextInput, MemoryInput, UrlsInput, UrlDataset  # noqa: F401
