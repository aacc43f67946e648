from typing import Any
ArrayLike = Any
