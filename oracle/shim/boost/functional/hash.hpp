// empty stand-in: the reference only needs this header to exist (TranscriptGroup.hpp:4)
