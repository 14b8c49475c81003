from .comparison import assert_close
