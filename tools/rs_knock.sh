# knock-out analysis of the row-streamed kernel (timings under a knock-out are not results): DLWPCS_RS_KNOCK bit 1 = no gathers,
# 2 = no MMAs, 4 = no epilogue math / stores; last line: the classic kernel on every layer (DLWPCS_RS=0)
for k in ${KNOCKS:-0 1 2 4 3 5 6 7}; do echo "knock $k"; DLWPCS_RS_KNOCK=$k timeout 100 python tools/layer_times.py; done
echo "classic"; DLWPCS_RS=0 timeout 100 python tools/layer_times.py
