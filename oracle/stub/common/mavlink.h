/*
 * oracle/stub/common/mavlink.h -- TEST INFRASTRUCTURE.  Stand-in for the MAVLink C headers that
 * /root/reference/uav_local_nav.c includes at :48 but does not vendor: just the types,
 * constants and call signatures that file names, with do-nothing inline bodies, so that the real translation
 * unit compiles and links against libuqs_mapping.so (oracle/link_reference.sh).  No MAVLink behaviour, wire
 * format or message id is claimed here beyond what the reference itself spells out; the linked program is
 * never run.
 */
#ifndef UQS_STUB_MAVLINK_H
#define UQS_STUB_MAVLINK_H
#include <stdint.h>
#include <string.h>

#define MAVLINK_MAX_PACKET_LEN 280
#define MAVLINK_COMM_0 0

typedef struct { uint32_t msgid; uint8_t sysid, compid, len; uint8_t payload[256]; } mavlink_message_t;
typedef struct { int parse_state; } mavlink_status_t;

enum { MAVLINK_MSG_ID_HEARTBEAT = 0, MAVLINK_MSG_ID_SYS_STATUS = 1, MAVLINK_MSG_ID_ATTITUDE = 30,
       MAVLINK_MSG_ID_LOCAL_POSITION_NED = 32, MAVLINK_MSG_ID_SERVO_OUTPUT_RAW = 36, MAVLINK_MSG_ID_COMMAND_ACK = 77,
       MAVLINK_MSG_ID_OPTICAL_FLOW = 100, MAVLINK_MSG_ID_OPTICAL_FLOW_RAD = 106, MAVLINK_MSG_ID_DISTANCE_SENSOR = 132,
       MAVLINK_MSG_ID_BATTERY_STATUS = 147, MAVLINK_MSG_ID_EXTENDED_SYS_STATE = 245, MAVLINK_MSG_ID_STATUSTEXT = 253 };
#define MAVLINK_MSG_SET_ATTITUDE_TARGET_FIELD_THRUST_BODY_LEN 3

enum { MAV_LANDED_STATE_UNDEFINED = 0, MAV_LANDED_STATE_ON_GROUND, MAV_LANDED_STATE_IN_AIR, MAV_LANDED_STATE_TAKEOFF,
       MAV_LANDED_STATE_LANDING };
enum { MAV_FRAME_LOCAL_NED = 1, MAV_FRAME_BODY_NED = 8, MAV_FRAME_BODY_OFFSET_NED = 9 };
enum { MAV_CMD_NAV_TAKEOFF = 22, MAV_CMD_DO_SET_MODE = 176, MAV_CMD_COMPONENT_ARM_DISARM = 400, MAV_CMD_SET_MESSAGE_INTERVAL = 511 };
enum { MAV_RESULT_ACCEPTED = 0, MAV_RESULT_TEMPORARILY_REJECTED = 1, MAV_RESULT_DENIED = 2 };
enum { MAV_MODE_FLAG_CUSTOM_MODE_ENABLED = 1, MAV_MODE_FLAG_SAFETY_ARMED = 128 };
enum { MAV_SYS_STATUS_SENSOR_3D_GYRO = 1, MAV_SYS_STATUS_SENSOR_Z_ALTITUDE_CONTROL = 0x2000,
       MAV_SYS_STATUS_SENSOR_XY_POSITION_CONTROL = 0x4000, MAV_SYS_STATUS_SENSOR_MOTOR_OUTPUTS = 0x8000 };
enum { MAV_TYPE_ONBOARD_CONTROLLER = 18, MAV_AUTOPILOT_INVALID = 8, MAV_STATE_ACTIVE = 4, MAV_COMP_ID_ONBOARD_COMPUTER = 191,
       MAV_DATA_STREAM_EXTRA3 = 12 };

typedef struct { uint32_t custom_mode; uint8_t type, autopilot, base_mode, system_status, mavlink_version; } mavlink_heartbeat_t;
typedef struct { uint16_t command; uint8_t result; } mavlink_command_ack_t;
typedef struct { uint8_t vtol_state, landed_state; } mavlink_extended_sys_state_t;
typedef struct { uint32_t onboard_control_sensors_present, onboard_control_sensors_enabled, onboard_control_sensors_health;
                 uint16_t load, voltage_battery; int16_t current_battery; int8_t battery_remaining; } mavlink_sys_status_t;
typedef struct { uint32_t time_usec; uint16_t servo1_raw, servo2_raw, servo3_raw, servo4_raw, servo5_raw, servo6_raw, servo7_raw,
                 servo8_raw; uint8_t port; } mavlink_servo_output_raw_t;
typedef struct { int32_t current_consumed, energy_consumed; int16_t temperature; uint16_t voltages[10]; int16_t current_battery;
                 uint8_t id, battery_function, type; int8_t battery_remaining; } mavlink_battery_status_t;
typedef struct { uint32_t time_boot_ms; float roll, pitch, yaw, rollspeed, pitchspeed, yawspeed; } mavlink_attitude_t;
typedef struct { uint64_t time_usec; float flow_comp_m_x, flow_comp_m_y, ground_distance; int16_t flow_x, flow_y; uint8_t sensor_id,
                 quality; float flow_rate_x, flow_rate_y; } mavlink_optical_flow_t;
typedef struct { uint64_t time_usec; uint32_t integration_time_us; float integrated_x, integrated_y, integrated_xgyro, integrated_ygyro,
                 integrated_zgyro; uint32_t time_delta_distance_us; float distance; int16_t temperature; uint8_t sensor_id, quality; }
        mavlink_optical_flow_rad_t;
typedef struct { uint32_t time_boot_ms; float x, y, z, vx, vy, vz; } mavlink_local_position_ned_t;
typedef struct { uint32_t time_boot_ms; uint16_t min_distance, max_distance, current_distance; uint8_t type, id, orientation,
                 covariance; } mavlink_distance_sensor_t;
typedef struct { uint8_t severity; char text[50]; } mavlink_statustext_t;
typedef struct { uint32_t time_boot_ms; float q[4]; float body_roll_rate, body_pitch_rate, body_yaw_rate, thrust; uint8_t target_system,
                 target_component, type_mask; float thrust_body[3]; } mavlink_set_attitude_target_t;
typedef struct { uint16_t chan1_raw, chan2_raw, chan3_raw, chan4_raw, chan5_raw, chan6_raw, chan7_raw, chan8_raw; uint8_t target_system,
                 target_component; } mavlink_rc_channels_override_t;

#define _MAV_PAYLOAD(msg) ((const char*)(&((msg)->payload[0])))
static inline uint8_t mavlink_parse_char(uint8_t chan, uint8_t c, mavlink_message_t* m, mavlink_status_t* s) { (void)chan; (void)c; (void)m; (void)s; return 0; }
static inline uint16_t mavlink_msg_to_send_buffer(uint8_t* buf, const mavlink_message_t* m) { (void)buf; (void)m; return 0; }

#define UQS_STUB_DECODE(name) \
  static inline void mavlink_msg_##name##_decode(const mavlink_message_t* m, mavlink_##name##_t* out) { (void)m; memset(out, 0, sizeof(*out)); }
UQS_STUB_DECODE(heartbeat) UQS_STUB_DECODE(command_ack) UQS_STUB_DECODE(extended_sys_state) UQS_STUB_DECODE(sys_status)
UQS_STUB_DECODE(servo_output_raw) UQS_STUB_DECODE(battery_status) UQS_STUB_DECODE(attitude) UQS_STUB_DECODE(optical_flow)
UQS_STUB_DECODE(optical_flow_rad) UQS_STUB_DECODE(local_position_ned) UQS_STUB_DECODE(distance_sensor) UQS_STUB_DECODE(statustext)

static inline uint16_t mavlink_msg_set_attitude_target_encode(uint8_t s, uint8_t c, mavlink_message_t* m, const mavlink_set_attitude_target_t* t) { (void)s; (void)c; (void)m; (void)t; return 0; }
static inline uint16_t mavlink_msg_rc_channels_override_encode(uint8_t s, uint8_t c, mavlink_message_t* m, const mavlink_rc_channels_override_t* t) { (void)s; (void)c; (void)m; (void)t; return 0; }
/* the *_pack calls: argument lists as the reference writes them (variadic here: only arity-agnostic acceptance is needed) */
static inline uint16_t uqs_stub_pack(uint8_t s, uint8_t c, mavlink_message_t* m, ...) { (void)s; (void)c; (void)m; return 0; }
#define mavlink_msg_command_long_pack uqs_stub_pack
#define mavlink_msg_request_data_stream_pack uqs_stub_pack
#define mavlink_msg_heartbeat_pack uqs_stub_pack
#define mavlink_msg_set_mode_pack uqs_stub_pack
#define mavlink_msg_set_position_target_local_ned_pack uqs_stub_pack
#endif
